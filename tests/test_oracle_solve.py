"""The banded restatement of the reference's f32 solver (oracle/ref_solve_band.c — CPU producer of
BASELINE config 5) against the reference's own f32 KATs, compared BITWISE like the reference's
`assert_eq!` on f32 (lib.rs:73-137, sparse.rs:1030-1080). With hb = n-1 the band covers the whole
matrix, so the routines run the reference's loops unabridged."""
import numpy as np

from oracle import ref_solve

f32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


def test_cholesky_decomposition_0():        # sparse.rs:1030-1059
    m = np.array([[4, 12, -16], [12, 37, -43], [-16, -43, 98]], f32)
    l = ref_solve.band_to_dense_lower(ref_solve.cholesky_band(ref_solve.dense_to_band(m, 2)))
    assert np.array_equal(bits(l), bits([[2, 0, 0], [6, 1, 0], [-8, 5, 3]]))


def test_cholesky_decomposition_1():        # sparse.rs:1061-1080
    m = np.array([[8, 0, 0, 0], [0, 7, 1, 0], [0, 1, 3, 0], [0, 0, 0, 2]], f32)
    l = ref_solve.band_to_dense_lower(ref_solve.cholesky_band(ref_solve.dense_to_band(m, 3)))
    ref = np.array([[2.828427, 0, 0, 0], [0, 2.6457512, 0, 0], [0, 0.37796451, 1.6903086, 0], [0, 0, 0, 1.4142135]], f32)
    assert np.array_equal(bits(l), bits(ref))


def test_forward_substitution_test_0():     # lib.rs:73-93
    l = np.array([[5, 0, 0], [8, 2, 0], [3, 7, 1]], f32)
    y = ref_solve.forward(ref_solve.dense_to_band(l, 2), np.array([7, 3, 1], f32))
    assert np.array_equal(bits(y), bits([f32(7.0) / f32(5.0), -4.1, 25.5]))


def test_backward_substitution_test_0():    # lib.rs:95-115
    l_star = np.array([[7, 1, 8], [0, 2, 3], [0, 0, 5]], f32)
    x = ref_solve.backward(ref_solve.dense_to_band(l_star.T.copy(), 2), np.array([1, 7, 3], f32))
    assert np.array_equal(bits(x), bits([f32(-32.0) / f32(35.0), 2.6, 0.6]))


def test_solve_test():                      # lib.rs:117-137
    a = np.array([[8, 0, 0, 0], [0, 7, 1, 0], [0, 1, 3, 0], [0, 0, 0, 2]], f32)
    x = ref_solve.solve_band(ref_solve.dense_to_band(a, 3), np.array([[5, 2, 8, 1]], f32))
    assert np.array_equal(bits(x[0]), bits([0.625, -0.1, 2.6999998, 0.5]))


def test_band_restriction_is_bit_equivalent():
    """A matrix of half-bandwidth 2 solved with hb = 2 and with hb = n-1 (the unabridged loops)."""
    rng = np.random.default_rng(0)
    n, hb = 40, 2
    a = np.zeros((n, n), f32)
    for i in range(n):
        for j in range(max(0, i - hb), i):
            a[i, j] = a[j, i] = f32(-rng.uniform(0.1, 1.0))
    a[np.arange(n), np.arange(n)] = f32(1.0) + np.abs(a).sum(axis=1).astype(f32)
    b = rng.uniform(0.5, 1.5, (3, n)).astype(f32)
    x_band = ref_solve.solve_band(ref_solve.dense_to_band(a, hb), b)
    x_full = ref_solve.solve_band(ref_solve.dense_to_band(a, n - 1), b)
    assert np.array_equal(bits(x_band), bits(x_full))
    assert np.abs(a.astype(np.float64) @ x_band.T.astype(np.float64) - b.T).max() < 1e-4


def test_spd_band_matches_generator():
    from basic_sparse_matrix_b200 import gen
    n, hb = 200, 5
    band = ref_solve.spd_band(n, hb)
    v, ci, ri, _ = gen.band(n, hb, dtype=f32)
    dense = np.zeros((n, n), f32)
    rows = np.repeat(np.arange(n), np.diff(ri.astype(np.int64)))
    dense[rows, ci.astype(np.int64)] = v
    assert np.array_equal(bits(ref_solve.band_to_dense_lower(band)), bits(np.tril(dense)))


def test_generic_csr_restatements_on_the_kats_and_against_the_band_oracle():
    """forward_csr / backward_csr (statement-by-statement restatements of lib.rs:28-65 on raw Csr parts — the checker of
    the GPU substitutions) reproduce the reference's KATs and the banded oracle bit for bit."""
    from basic_sparse_matrix_b200 import Csr
    l = Csr.from_data([[5, 0, 0], [8, 2, 0], [3, 7, 1]], f32)
    y = ref_solve.forward_csr(*l.raw_parts(), np.array([[7, 3, 1]], f32))
    assert np.array_equal(bits(y[0]), bits([f32(7.0) / f32(5.0), -4.1, 25.5]))                       # lib.rs:73-93
    ls = Csr.from_data([[7, 1, 8], [0, 2, 3], [0, 0, 5]], f32)
    x = ref_solve.backward_csr(*ls.raw_parts(), np.array([[1, 7, 3]], f32))
    assert np.array_equal(bits(x[0]), bits([f32(-32.0) / f32(35.0), 2.6, 0.6]))                      # lib.rs:95-115
    n, hb = 60, 4
    a_band = ref_solve.spd_band(n, hb)
    b = np.random.default_rng(1).uniform(0.5, 1.5, (3, n)).astype(f32)
    l_band = ref_solve.cholesky_band(a_band)
    y = ref_solve.forward_csr(*ref_solve.band_to_csr_lower(l_band), b)
    x = ref_solve.backward_csr(*ref_solve.band_to_csr_upper(l_band), y)
    assert np.array_equal(bits(x), bits(ref_solve.solve_band(a_band, b)))
