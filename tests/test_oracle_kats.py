"""CPU: pin the oracle on every golden vector the reference's own tests hold for this path
(tests/golden/reference_kats.json, transcribed from /root/reference/src/sparse.rs and dense.rs)."""
import numpy as np
import pytest

from helpers import assert_bitwise, random_csr, random_dense
from oracle import ref_numpy
from oracle.ref_cpu import OracleCsr


def test_structure_kats(golden):
    # example_mat_0..2 (sparse.rs:815-852), csr_with_empty_row_top/middle (1111-1151)
    for k in golden["structure"]:
        if "data" not in k:
            continue
        m = OracleCsr.from_data(k["data"])
        v, ci, ri = m.raw()
        assert v.tolist() == k["v"], k["name"]
        assert ci.tolist() == k["col_index"], k["name"]
        assert ri.tolist() == k["row_index"], k["name"]
        assert m.is_finalised


def test_create_mat_by_insert(golden):
    k = [s for s in golden["structure"] if s["name"] == "create_mat_by_insert"][0]   # sparse.rs:854-868
    b = OracleCsr.new(tuple(k["dims"]))
    for v, r, c in k["inserts"]:
        b.insert(v, r, c)
    b.finalise()
    ref = OracleCsr.from_data(k["equals_from_data"])
    for x, y in zip(b.raw(), ref.raw()):
        assert x.tolist() == y.tolist()


@pytest.mark.parametrize("dtype", [np.int32, np.float32, np.float64])
def test_mul_dense_kats(golden, dtype):
    # test_dense_mul (sparse.rs:1082-1109) and test_nnz (1153-1178); values are small integers so
    # the float instantiations must reproduce them exactly too
    for k in golden["mul_dense"]:
        a = OracleCsr.from_data(k["csr_rows"], dtype)
        out = a.mul_dense([np.array(c, dtype) for c in k["dense_columns"]])
        ref = OracleCsr.from_data(k["output_rows"], dtype)
        for x, y in zip(out.raw(), ref.raw()):
            assert x.tolist() == y.tolist(), k["name"]
        assert out.is_finalised
        if "nnz" in k:
            assert out.get_nnz() == k["nnz"]


def test_expected_raw_arrays():
    # raw layout of the two results
    a = OracleCsr.from_data([[3, 0, 2, 0], [7, 0, 0, 0], [0, 2, 0, 1], [0, 0, 1, 0], [1, 0, 0, 0]])
    out = a.mul_dense([np.array(c, np.int32) for c in ([1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12])])
    assert out.v.tolist() == [9, 29, 49, 7, 35, 63, 8, 20, 32, 3, 7, 11, 1, 5, 9]
    assert out.col_index.tolist() == [0, 1, 2] * 5
    assert out.row_index.tolist() == [0, 3, 6, 9, 12, 15]
    m = OracleCsr.from_data([[5, 2, 1, 3], [7, 0, 1, 3], [0, 1, 0, 0], [0, 7, 4, 0]])
    out = m.mul_dense([np.array(c, np.int32) for c in ([1, 0, 3, 4], [8, 0, 0, 5])])
    assert out.v.tolist() == [20, 55, 22, 71, 12]
    assert out.col_index.tolist() == [0, 1, 0, 1, 0]
    assert out.row_index.tolist() == [0, 2, 4, 4, 5]


def test_mul_dense_incorrect_dimensions():
    a = OracleCsr.from_data([[1, 2, 3]])
    with pytest.raises(RuntimeError, match="IncorrectDimensions"):        # sparse.rs:427-429
        a.mul_dense([np.array([1, 2], np.int32)])


def test_mul_vector_kat(golden):
    k = golden["mul_vector"]                                               # sparse.rs:1501-1529
    with pytest.raises(RuntimeError, match="IncorrectDimensions"):
        OracleCsr.from_data(k["bad_dims_matrix"]).mul_vector(k["v"], out_len=5)
    assert OracleCsr.from_data(k["identity"]).mul_vector(k["v"]).tolist() == k["v"]
    assert OracleCsr.from_data(k["matrix"]).mul_vector(k["v"]).tolist() == k["expected"]


def test_insert_after_finalise_and_zero_skip():
    m = OracleCsr.new((2, 2), np.float64)
    m.insert(0.0, 0, 0)       # skipped
    m.insert(-0.0, 0, 1)      # -0.0 == 0.0 -> skipped
    m.insert(float("nan"), 1, 0)   # NaN != 0 -> kept
    m.finalise()
    assert m.get_nnz() == 1 and np.isnan(m.v[0]) and m.row_index.tolist() == [0, 0, 1]
    with pytest.raises(RuntimeError, match="MatrixFinalised"):             # sparse.rs:223-225
        m.insert(1.0, 1, 1)


def test_out_of_order_inserts_pile_into_last_row():
    # the reference never re-sorts (sparse.rs:237-250): the bench relies on this
    m = OracleCsr.new((4, 4), np.int32)
    for v, r, c in [(1, 2, 1), (2, 0, 3), (3, 3, 0), (4, 1, 1)]:
        m.insert(v, r, c)
    m.finalise()
    assert m.v.tolist() == [1, 2, 3, 4]
    assert m.col_index.tolist() == [1, 3, 0, 1]
    assert m.row_index.tolist() == [0, 0, 0, 2, 4]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_numpy_restatement_matches_c_oracle(dtype):
    """ref_numpy (vectorised across rows) must agree bit-for-bit with the pinned C restatement,
    including unsorted / duplicate columns, empty rows, and the zero-dropped result Csr."""
    rng = np.random.default_rng(7)
    for (m, k, n) in [(37, 29, 1), (64, 50, 10), (33, 40, 64)]:
        v, ci, ri = random_csr(rng, m, k, dtype, giant_row=5, giant_len=300)
        b = random_dense(rng, k, n, dtype)
        b[3, :] = 0.0
        a = OracleCsr.from_raw((m, k), v, ci, ri)
        want = a.mul_dense_rows([b[:, c].copy() for c in range(n)], 0, m)
        got = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
        assert_bitwise(got, want, f"{m}x{k}x{n}")
        res = a.mul_dense([b[:, c].copy() for c in range(n)])
        rv, rc, rr = ref_numpy.dense_to_csr(got)
        assert_bitwise(rv, res.v)
        assert rc.tolist() == res.col_index.tolist() and rr.tolist() == res.row_index.tolist()


def test_faithful_and_lean_agree():
    rng = np.random.default_rng(3)
    v, ci, ri = random_csr(rng, 50, 60, np.float64)
    b = random_dense(rng, 60, 4, np.float64)
    a = OracleCsr.from_raw((50, 60), v, ci, ri)
    cols = [b[:, c].copy() for c in range(4)]
    x, y = a.mul_dense(cols, faithful=True), a.mul_dense(cols, faithful=False)
    for p, q in zip(x.raw(), y.raw()):
        assert np.array_equal(p, q)
