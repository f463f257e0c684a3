"""Shared builders for the parity tests (test infrastructure)."""
import numpy as np


def random_csr(rng, rows, cols, dtype, mean_len=6.0, empty_frac=0.2, giant_row=None, giant_len=0,
               exact=False, sorted_cols=False):
    """Adversarial CSR in the reference layout: empty rows, unsorted and duplicate columns inside
    a row (legal: src/sparse.rs:237-250 never sorts), optional giant row.
    exact=True draws dyadic values k/8 so every product and partial sum is exact."""
    lens = rng.poisson(mean_len, size=rows).astype(np.int64)
    lens[rng.random(rows) < empty_frac] = 0
    if giant_row is not None:
        lens[giant_row] = giant_len
    row_index = np.zeros(rows + 1, dtype=np.uint64)
    np.cumsum(lens, out=row_index[1:])
    nnz = int(row_index[-1])
    col_index = rng.integers(0, cols, size=nnz).astype(np.uint64)
    if sorted_cols:
        for r in range(rows):
            s, e = int(row_index[r]), int(row_index[r + 1])
            col_index[s:e] = np.sort(col_index[s:e])
    if exact:
        v = (rng.integers(-16, 17, size=nnz).astype(np.float64) / 8.0)
        v[v == 0] = 0.125
    else:
        v = rng.standard_normal(nnz)
    return v.astype(dtype), col_index, row_index


def random_dense(rng, rows, cols, dtype, exact=False):
    if exact:
        return (rng.integers(-32, 33, size=(rows, cols)).astype(np.float64) / 16.0).astype(dtype)
    return rng.standard_normal((rows, cols)).astype(dtype)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64 if a.dtype == np.float64 else np.uint32)


def assert_bitwise(got, want, what=""):
    got = np.ascontiguousarray(got)
    want = np.ascontiguousarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    # -0.0 vs +0.0 are both "zero" to the reference (dropped by insert); compare them as equal
    g = np.where(got == 0, 0.0, got).astype(got.dtype)
    w = np.where(want == 0, 0.0, want).astype(want.dtype)
    bad = bits(g) != bits(w)
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} elements differ bitwise; first at "
                             f"{idx.tolist()} got {got[tuple(idx[0])]!r} want {want[tuple(idx[0])]!r}")


def assert_tolerance(got, want, scale, tol, what=""):
    """|got - want| <= tol * sum_j |a_ij b_jk|  (the stated metric; equals the plain
    relative error when nothing cancels)."""
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bound = tol * np.maximum(scale.astype(np.float64), np.finfo(np.float64).tiny)
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err / bound), err.shape)
        raise AssertionError(f"{what}: {bad.sum()} elements exceed tol {tol}; worst at {i}: got {got[i]!r} "
                             f"want {want[i]!r} err/scale {err[i] / max(scale[i], 1e-300):.3e}")
