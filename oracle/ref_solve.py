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
