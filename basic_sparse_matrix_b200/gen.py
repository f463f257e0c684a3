"""Synthetic workloads — numpy mirror of csrc/gen.cu (bench / test support).

Every generator is a pure function of ``(seed, logical index)`` through ``hash_u64``, so the
device-side generators (``gpu.DeviceCsr.laplacian/band/rmat``, ``gpu.DeviceDense.generate``) and
this module produce identical arrays, and any row of a full-size operand can be regenerated on the
host without moving it over PCIe.

Shapes follow BASELINE.json's configs: the reference bench as written
(/root/reference/benches/sparse_dense_mul.rs:6-35), 2-D / 3-D Laplacians, R-MAT, SPD band.
The reference draws from rand 0.8.5's ``StdRng`` (ChaCha12), which cannot be reproduced without a
Rust toolchain; the stream below is ours and documented.
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1
_U = np.uint64

MODE_EXACT, MODE_REAL, MODE_EXACT_SMALL = 0, 1, 2


def _mix64(z):
    z = (z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)
    return z ^ (z >> _U(31))


def hash_u64(seed: int, i) -> np.ndarray:
    """h(seed, i) = mix64((i+1)*0x9E3779B97F4A7C15 + seed*0xD1342543DE82EF95) mod 2^64 (gen.cu)."""
    i = np.asarray(i, dtype=np.uint64)
    s = _U((int(seed) * 0xD1342543DE82EF95) & _M64)
    with np.errstate(over="ignore"):
        return _mix64((i + _U(1)) * _U(0x9E3779B97F4A7C15) + s)


def _unit(h):
    return (h >> _U(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def dense_value(h, mode: int, offset: float = 0.0):
    m = h >> _U(11)
    if mode == MODE_EXACT:
        return (m % _U(1024)).astype(np.float64) / 1024.0 + offset
    if mode == MODE_EXACT_SMALL:
        return (m % _U(16)).astype(np.float64) / 16.0 + offset
    return m.astype(np.float64) * (1.0 / 9007199254740992.0) + offset


def sparse_value(h, mode: int):
    m = h >> _U(11)
    if mode == MODE_EXACT:
        return (_U(1) + m % _U(256)).astype(np.float64) / 256.0
    if mode == MODE_EXACT_SMALL:
        return (_U(1) + m % _U(8)).astype(np.float64) / 8.0
    return 0.5 + m.astype(np.float64) * (1.0 / 9007199254740992.0)


def dense_rows(rows_total: int, cols: int, seed: int, mode: int, offset: float = 0.0, dtype=np.float64,
               row_ids=None) -> np.ndarray:
    """Row-major block of the generated dense operand: all rows, or the given row ids."""
    r = np.arange(rows_total, dtype=np.uint64) if row_ids is None else np.asarray(row_ids, dtype=np.uint64)
    idx = r[:, None] * _U(cols) + np.arange(cols, dtype=np.uint64)[None, :]
    return dense_value(hash_u64(seed, idx), mode, offset).astype(dtype)


# ---- Laplacians (configs 2 and 4) -------------------------------------------------------------
def laplacian_row_counts(nx, ny, nz, row_begin=0, row_end=None) -> np.ndarray:
    n = nx * ny * nz
    row_end = n if row_end is None else row_end
    i = np.arange(row_begin, row_end, dtype=np.int64)
    x, y, z = i % nx, (i // nx) % ny, i // (nx * ny)
    return (1 + (x > 0) + (x + 1 < nx) + (y > 0) + (y + 1 < ny) + (z > 0) + (z + 1 < nz)).astype(np.int64)


def laplacian(nx, ny, nz=1, row_begin=0, row_end=None, dtype=np.float64):
    """5-point (nz=1) / 7-point Laplacian rows [row_begin,row_end): diagonal 2*dims, off-diagonals
    -1, columns ascending. Returns (v, col_index[u64], row_index[u64], dims)."""
    n = nx * ny * nz
    row_end = n if row_end is None else row_end
    i = np.arange(row_begin, row_end, dtype=np.int64)
    x, y, z = i % nx, (i // nx) % ny, i // (nx * ny)
    diag = 2.0 * ((nx > 1) + (ny > 1) + (nz > 1))
    cand_cols = np.stack([i - nx * ny, i - nx, i - 1, i, i + 1, i + nx, i + nx * ny], axis=1)
    cand_ok = np.stack([z > 0, y > 0, x > 0, np.ones_like(x, bool), x + 1 < nx, y + 1 < ny, z + 1 < nz], axis=1)
    cand_val = np.broadcast_to(np.array([-1, -1, -1, diag, -1, -1, -1], dtype=np.float64), cand_cols.shape)
    counts = cand_ok.sum(axis=1)
    row_index = np.zeros(len(i) + 1, dtype=np.uint64)
    np.cumsum(counts, out=row_index[1:])
    return (cand_val[cand_ok].astype(dtype), cand_cols[cand_ok].astype(np.uint64), row_index, (len(i), n))


# ---- SPD band (config 5) ----------------------------------------------------------------------------
def band(n, hb, row_begin=0, row_end=None, dtype=np.float64):
    """a_ij = -1/(1+|i-j|) for 0<|i-j|<=hb, a_ii = 1 + sum_j |a_ij| (summed over ascending j in f64)."""
    row_end = n if row_end is None else row_end
    i = np.arange(row_begin, row_end, dtype=np.int64)
    lo = np.minimum(i, hb)
    hi = np.minimum(n - 1 - i, hb)
    diag = np.ones(len(i), dtype=np.float64)
    for d in range(hb, 0, -1):
        diag[lo >= d] += 1.0 / (1.0 + d)
    for d in range(1, hb + 1):
        diag[hi >= d] += 1.0 / (1.0 + d)
    offs = np.arange(-hb, hb + 1, dtype=np.int64)
    cols = i[:, None] + offs[None, :]
    ok = (offs[None, :] >= -lo[:, None]) & (offs[None, :] <= hi[:, None])
    vals = np.where(offs[None, :] == 0, diag[:, None], -1.0 / (1.0 + np.abs(offs[None, :]).astype(np.float64)))
    counts = ok.sum(axis=1)
    row_index = np.zeros(len(i) + 1, dtype=np.uint64)
    np.cumsum(counts, out=row_index[1:])
    return vals[ok].astype(dtype), cols[ok].astype(np.uint64), row_index, (len(i), n)


# ---- R-MAT (config 3) ---------------------------------------------------------------------------------
def rmat(scale, edges, a=0.57, b=0.19, c=0.19, seed=3, mode=MODE_EXACT, dtype=np.float64):
    """2^scale square R-MAT: `edges` draws, sorted by (row,col), duplicates KEPT (legal in the
    reference: insert_unchecked never dedups, src/sparse.rs:237-250)."""
    e = np.arange(edges, dtype=np.uint64)
    row = np.zeros(edges, dtype=np.uint64)
    col = np.zeros(edges, dtype=np.uint64)
    ab, abc = a + b, a + b + c
    for l in range(scale):
        u = _unit(hash_u64(seed, e * _U(64) + _U(l)))
        rbit = u >= ab
        cbit = ((u >= a) & (u < ab)) | (u >= abc)
        row = (row << _U(1)) | rbit.astype(np.uint64)
        col = (col << _U(1)) | cbit.astype(np.uint64)
    keys = np.sort((row << _U(32)) | col)
    rows_n = 1 << scale
    row_index = np.searchsorted(keys, np.arange(rows_n + 1, dtype=np.uint64) << _U(32), side="left").astype(np.uint64)
    col_index = keys & _U(0xFFFFFFFF)
    vals = sparse_value(hash_u64(seed + 1, e), mode).astype(dtype)
    return vals, col_index, row_index, (rows_n, rows_n)


# ---- the reference bench as written (config 1) -------------------------------------------------------
def bench_as_written(e: int, seed: int = 1000, dtype=np.float64, sort_inserts: bool = False):
    """``sd_mul`` of /root/reference/benches/sparse_dense_mul.rs:6-35 with f64 for u32:
    A = 1000x1000 ``Csr::new``, ``e`` draws (row=r%1000, col=r%1000, v=r%255) pushed through
    ``insert`` IN DRAW ORDER (zeros skipped; out-of-order rows pile up in the last row — the
    reference never re-sorts, sparse.rs:237-250); x = Dense 10 cols x 1000 rows with e/100 cells
    assigned.  Returns (Csr, Dense) built with this package's host types.
    ``sort_inserts=True`` is the as-intended variant (draws sorted by (row,col))."""
    from .dense import Dense
    from .sparse import Csr
    k = np.arange(3 * e + 3 * (e // 100), dtype=np.uint64)
    r = hash_u64(seed, k)
    rows = (r[0:3 * e:3] % _U(1000)).astype(np.int64)
    cols = (r[1:3 * e:3] % _U(1000)).astype(np.int64)
    vals = (r[2:3 * e:3] % _U(255)).astype(np.float64)
    if sort_inserts:
        order = np.lexsort((cols, rows))
        rows, cols, vals = rows[order], cols[order], vals[order]
    a = Csr.new((1000, 1000), dtype)
    for rr, cc, vv in zip(rows.tolist(), cols.tolist(), vals.tolist()):
        a.insert(vv, rr, cc)
    a = a.finalise()
    x = Dense.new_default_with_dims(10, 1000, dtype)
    base = 3 * e
    m = e // 100
    xc = (r[base:base + 3 * m:3] % _U(10)).astype(np.int64)
    xr = (r[base + 1:base + 3 * m:3] % _U(1000)).astype(np.int64)
    xv = (r[base + 2:base + 3 * m:3] % _U(255)).astype(np.float64)
    for c_, r_, v_ in zip(xc.tolist(), xr.tolist(), xv.tolist()):
        x.get_col_mut(c_)[r_] = v_
    return a, x
