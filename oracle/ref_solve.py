"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/ref_solve_band.c: the banded restatement
of the reference's f32 `solve` (lib.rs:11-65, sparse.rs:682-714), the CPU producer of BASELINE config 5."""
import ctypes as C

import numpy as np

from . import ref_cpu

_F = C.POINTER(C.c_float)


def _p(a):
    return a.ctypes.data_as(_F)


def dense_to_band(a: np.ndarray, hb: int) -> np.ndarray:
    """Lower band of a square matrix in the oracle's band storage: band[i, j-i+hb] = a[i, j]."""
    n = a.shape[0]
    band = np.zeros((n, hb + 1), np.float32)
    for i in range(n):
        for j in range(max(0, i - hb), i + 1):
            band[i, j - i + hb] = a[i, j]
    return band


def band_to_dense_lower(band: np.ndarray) -> np.ndarray:
    n, w = band.shape
    hb = w - 1
    out = np.zeros((n, n), np.float32)
    for i in range(n):
        for j in range(max(0, i - hb), i + 1):
            out[i, j] = band[i, j - i + hb]
    return out


def spd_band(n: int, hb: int) -> np.ndarray:
    """Band storage of BASELINE config 5's matrix (same formula as gen.band / bsm_gen_band), f32:
    a_ij = -1/(1+|i-j|) inside the band, a_ii = 1 + sum_j |a_ij|."""
    band = np.zeros((n, hb + 1), np.float32)
    d = np.arange(hb, 0, -1, dtype=np.float64)           # |i-j| for slots 0..hb-1
    off = (-1.0 / (1.0 + d)).astype(np.float32)
    i = np.arange(n)
    for s in range(hb):
        band[:, s] = np.where(i - (hb - s) >= 0, off[s], 0.0)
    from basic_sparse_matrix_b200 import gen
    v, ci, ri, _ = gen.band(n, hb, dtype=np.float32)
    diag = v[(ci == np.repeat(np.arange(n, dtype=np.uint64), np.diff(ri.astype(np.int64))))]
    band[:, hb] = diag
    return band


def cholesky_band(a_band: np.ndarray) -> np.ndarray:
    a_band = np.ascontiguousarray(a_band, np.float32)
    n, w = a_band.shape
    l_band = np.zeros_like(a_band)
    L = ref_cpu.lib()
    L.osolve_cholesky_band_f32.restype = C.c_int
    L.osolve_cholesky_band_f32.argtypes = [C.c_size_t, C.c_size_t, _F, _F]
    assert L.osolve_cholesky_band_f32(n, w - 1, _p(a_band), _p(l_band)) == 0
    return l_band


def forward(l_band: np.ndarray, b: np.ndarray) -> np.ndarray:
    l_band = np.ascontiguousarray(l_band, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    y = np.zeros_like(b)
    L = ref_cpu.lib()
    L.osolve_forward_band_f32.restype = None
    L.osolve_forward_band_f32.argtypes = [C.c_size_t, C.c_size_t, _F, _F, _F]
    L.osolve_forward_band_f32(l_band.shape[0], l_band.shape[1] - 1, _p(l_band), _p(b), _p(y))
    return y


def backward(l_band: np.ndarray, y: np.ndarray) -> np.ndarray:
    l_band = np.ascontiguousarray(l_band, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    x = np.zeros_like(y)
    L = ref_cpu.lib()
    L.osolve_backward_band_f32.restype = None
    L.osolve_backward_band_f32.argtypes = [C.c_size_t, C.c_size_t, _F, _F, _F]
    L.osolve_backward_band_f32(l_band.shape[0], l_band.shape[1] - 1, _p(l_band), _p(y), _p(x))
    return x


def solve_band(a_band: np.ndarray, b_cols: np.ndarray) -> np.ndarray:
    """`solve(a, b)` for nrhs right-hand sides; b_cols is (nrhs, n): one row per COLUMN of the reference's
    Dense (column-major). Returns x in the same layout."""
    a_band = np.ascontiguousarray(a_band, np.float32)
    b_cols = np.ascontiguousarray(b_cols, np.float32)
    n, w = a_band.shape
    nrhs = b_cols.shape[0]
    x = np.zeros_like(b_cols)
    l_band = np.zeros_like(a_band)
    y = np.zeros(n, np.float32)
    L = ref_cpu.lib()
    L.osolve_band_f32.restype = C.c_int
    L.osolve_band_f32.argtypes = [C.c_size_t, C.c_size_t, _F, C.c_size_t, _F, _F, _F, _F]
    assert L.osolve_band_f32(n, w - 1, _p(a_band), nrhs, _p(b_cols), _p(x), _p(l_band), _p(y)) == 0
    return x


# ---- generic CSR restatements (small cases; pure Python loops over the reference's own statements) ----------------------
def forward_csr(v, ci, ri, b_cols):
    """lib.rs:28-46 on a Csr given by its raw parts; b_cols is (nrhs, n), one row per Dense column. f32 or f64."""
    dt = v.dtype.type
    nrhs, n = b_cols.shape
    y = np.zeros_like(b_cols)
    for c in range(nrhs):                                       # :32
        for r in range(n):                                      # :33
            l_x = dt(0.0)                                       # :35
            s, e = int(ri[r]), int(ri[r + 1])
            for k in range(s, e):                               # :37
                if int(ci[k]) != r:                             # :38
                    l_x = dt(l_x + dt(v[k] * y[c, int(ci[k])]))  # :39
            y[c, r] = dt(dt(b_cols[c, r] - l_x) / v[e - 1])     # :42  row.last()
    return y


def backward_csr(v, ci, ri, y_cols):
    """lib.rs:49-65 on a Csr (l_star) given by its raw parts."""
    dt = v.dtype.type
    nrhs, n = y_cols.shape
    x = np.zeros_like(y_cols)
    for c in range(nrhs):                                       # :53
        for r in range(n - 1, -1, -1):                          # :54
            l_x = dt(0.0)                                       # :56
            s, e = int(ri[r]), int(ri[r + 1])
            for k in range(s + 1, e):                           # :57  skip(1)
                l_x = dt(l_x + dt(v[k] * x[c, int(ci[k])]))     # :58
            x[c, r] = dt(dt(y_cols[c, r] - l_x) / v[s])         # :60  row[0]
    return x


def band_to_csr_lower(l_band: np.ndarray):
    """Csr raw parts of the lower-triangular factor held in band storage: what `l.insert` builds in cholesky_decomp
    (sparse.rs:710 -> 229: zeros are not stored), rows ascending, columns ascending (diagonal last)."""
    n, w = l_band.shape
    hb = w - 1
    i = np.arange(n)[:, None]
    j = i - hb + np.arange(w)[None, :]
    keep = (j >= 0) & (l_band != 0)
    ri = np.zeros(n + 1, np.uint64)
    np.cumsum(keep.sum(axis=1), out=ri[1:])
    return l_band[keep].astype(l_band.dtype), j[keep].astype(np.uint64), ri


def band_to_csr_upper(l_band: np.ndarray):
    """Csr raw parts of l.transpose() (sparse.rs:296-318): row r holds L[j][r] for j = r .. r+hb, columns ascending
    (diagonal first), zeros not stored."""
    n, w = l_band.shape
    hb = w - 1
    # element (r, j) of the transpose = L[j][r] = band[j, r - j + hb], j in [r, r + hb]
    r = np.arange(n)[:, None]
    j = r + np.arange(w)[None, :]
    ok = j < n
    jj = np.where(ok, j, 0)
    vals = np.where(ok, l_band[jj, (r - jj + hb) % w], 0)
    keep = ok & (vals != 0)
    ri = np.zeros(n + 1, np.uint64)
    np.cumsum(keep.sum(axis=1), out=ri[1:])
    return vals[keep].astype(l_band.dtype), j[keep].astype(np.uint64), ri
