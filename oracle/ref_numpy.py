"""ORACLE — TEST INFRASTRUCTURE ONLY (not the product path).

numpy restatement of the arithmetic of ``Csr<T>::mul_dense``
(/root/reference/src/sparse.rs:426-446) for mid-size parity checks: the per-row sum is taken
over the stored entries in stored order, ``value = value + (a*b)`` from ``T::default()``
(:434-439), multiply and add rounded separately (numpy never fuses them).  Vectorised ACROSS
rows, sequential WITHIN a row, so every output element sees exactly the reference's order.

``dense_to_csr`` restates the result construction: every output goes through ``insert``
(:442 -> :222-233, values equal to ``T::default()`` skipped; ``-0.0`` skipped, ``NaN`` kept) and
``finalise`` (:206-219).

Parity status: validated against the pinned C oracle (tests/test_oracle_kats.py).
"""
from __future__ import annotations

import numpy as np


def mul_dense_rowmajor(v, col_index, row_index, b_rowmajor, row_begin=0, row_end=None):
    """Dense row-major C[row_begin:row_end] = A[row_begin:row_end] @ B in the reference's
    summation order. ``b_rowmajor`` is k x n (the oracle does not care about B's layout, only
    about which element is read: B[col_index[e], c] == rhs.get_col(c)[col_index[e]], :437)."""
    v = np.asarray(v)
    col_index = np.asarray(col_index).astype(np.int64)
    row_index = np.asarray(row_index).astype(np.int64)
    b = np.asarray(b_rowmajor)
    if b.ndim == 1:
        b = b[:, None]
    m = row_index.shape[0] - 1
    row_end = m if row_end is None else row_end
    starts = row_index[row_begin:row_end]
    lens = row_index[row_begin + 1:row_end + 1] - starts
    acc = np.zeros((row_end - row_begin, b.shape[1]), dtype=v.dtype)   # T::default()  :434
    max_len = int(lens.max()) if lens.size else 0
    for j in range(max_len):                                           # stored order  :435
        sel = np.nonzero(lens > j)[0]
        e = starts[sel] + j
        c = v[e][:, None] * b[col_index[e]]                            # c = a*b       :438
        acc[sel] = acc[sel] + c                                        # value + c     :439
    return acc


def dense_to_csr(dense_rowmajor):
    """Zero-dropping result construction (:442 + :222-250 + :206-219) for in-order inserts:
    returns (v, col_index[u64], row_index[u64] of length rows+1)."""
    d = np.asarray(dense_rowmajor)
    keep = d != 0          # value != T::default(): NaN kept, -0.0 dropped   :229
    counts = keep.sum(axis=1).astype(np.uint64)
    row_index = np.zeros(d.shape[0] + 1, np.uint64)
    np.cumsum(counts, out=row_index[1:])
    rr, cc = np.nonzero(keep)      # row-major order == insertion order (:431,:433)
    return d[rr, cc], cc.astype(np.uint64), row_index


def abs_product_sum(v, col_index, row_index, b_rowmajor):
    """sum_j |a_ij * b_jk| — the denominator of the stated tolerance metric."""
    return mul_dense_rowmajor(np.abs(v), col_index, row_index, np.abs(b_rowmajor))
