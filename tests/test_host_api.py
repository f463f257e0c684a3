"""CPU: the host-side mirror of the reference interface (Csr / Dense / MatDim / MatErr) and the
loud failure of the product path when no CUDA device is usable (no CPU fallback)."""
import numpy as np
import pytest

from basic_sparse_matrix_b200 import Csr, Dense, DenseS, MatDim, MatErr, MatError
from basic_sparse_matrix_b200 import _lib


def test_structure_kats(golden):
    for k in golden["structure"]:
        if "data" not in k:
            continue
        m = Csr.from_data(k["data"])
        assert m.v.tolist() == k["v"] and m.col_index.tolist() == k["col_index"]
        assert m.row_index.tolist() == k["row_index"] and m.is_finalised


def test_create_mat_by_insert(golden):
    k = [s for s in golden["structure"] if s["name"] == "create_mat_by_insert"][0]
    b = Csr.new(tuple(k["dims"]))
    for v, r, c in k["inserts"]:
        b.insert(v, r, c)
    assert b.finalise() == Csr.from_data(k["equals_from_data"])


def test_dense_kats(golden):
    init, get_col = golden["dense"]
    a = Dense.new_default_with_dims(*init["new_default_with_dims"])       # (cols, rows)  dense.rs:13
    assert a == Dense.from_data(init["equals_from_data_columns"])
    assert a.get_dims() == MatDim(rows=7, cols=5)
    d = Dense.from_data(get_col["columns"])
    assert d.get_col(2).tolist() == get_col["get_col_2"]                   # dense.rs:82-90


def test_insert_rules():
    m = Csr.new((2, 2))
    m.insert(0.0, 0, 0)
    m.insert(-0.0, 0, 1)
    m.insert(float("nan"), 1, 0)
    m = m.finalise()
    assert m.get_nnz() == 1 and m.row_index.tolist() == [0, 0, 1]
    with pytest.raises(MatError) as e:
        m.insert(1.0, 1, 1)
    assert e.value.kind == MatErr.MatrixFinalised
    # out-of-order rows are appended to the current last row
    m = Csr.new((4, 4))
    for v, r, c in [(1, 2, 1), (2, 0, 3), (3, 3, 0), (4, 1, 1)]:
        m.insert(v, r, c)
    m = m.finalise()
    assert m.col_index.tolist() == [1, 3, 0, 1] and m.row_index.tolist() == [0, 0, 0, 2, 4]


def test_host_csr_matches_oracle_construction():
    from oracle.ref_cpu import OracleCsr
    rng = np.random.default_rng(0)
    data = rng.integers(-2, 3, size=(9, 7)).astype(np.float64)
    h, o = Csr.from_data(data.tolist()), OracleCsr.from_data(data, np.float64)
    assert np.array_equal(h.v, o.v) and np.array_equal(h.col_index, o.col_index)
    assert np.array_equal(h.row_index, o.row_index)
    # random (possibly out-of-order) inserts
    h, o = Csr.new((6, 6)), OracleCsr.new((6, 6), np.float64)
    for _ in range(40):
        v, r, c = float(rng.integers(0, 3)), int(rng.integers(0, 6)), int(rng.integers(0, 6))
        h.insert(v, r, c)
        o.insert(v, r, c)
    h.finalise()
    o.finalise()
    assert np.array_equal(h.v, o.v) and np.array_equal(h.col_index, o.col_index)
    assert np.array_equal(h.row_index, o.row_index)


def test_insert_sequences_match_the_oracle_property():
    """Any sequence of inserts — rows out of order, gaps, zeros, rows beyond the declared dimensions — builds the same three arrays in
    the Python mirror as in the C oracle (sparse.rs:222-250, 206-219), or fails the same way (finalise's panic, 212-214)."""
    from hypothesis import given, settings, strategies as st
    from oracle.ref_cpu import OracleCsr

    entry = st.tuples(st.sampled_from([0.0, -0.0, 1.0, -2.5, 3.0, float("inf")]), st.integers(0, 9), st.integers(0, 7))

    @settings(max_examples=200, deadline=None)
    @given(rows=st.integers(1, 8), entries=st.lists(entry, max_size=24))
    def check(rows, entries):
        h, o = Csr.new((rows, 8)), OracleCsr.new((rows, 8), np.float64)
        for v, r, c in entries:
            h.insert(v, r, c)
            o.insert(v, r, c)
        try:
            o.finalise()
        except RuntimeError as e:
            assert "big eek" in str(e)
            with pytest.raises(RuntimeError):
                h.finalise()
            return
        h.finalise()
        assert np.array_equal(h.v.view(np.uint64), o.v.view(np.uint64))
        assert np.array_equal(h.col_index, o.col_index) and np.array_equal(h.row_index, o.row_index)
        assert h.get_nnz() == o.get_nnz()
        with pytest.raises(MatError):
            h.insert(1.0, 0, 0)                                              # MatrixFinalised on both sides
        with pytest.raises(RuntimeError):
            o.insert(1.0, 0, 0)

    check()


def test_get_row_compact():
    m = Csr.from_data([[8, 0, 2, 0, 0], [0, 0, 5, 0, 0], [0, 0, 0, 0, 0]])
    row = m.get_row_compact(0)
    assert [(e.v, e.col_index, e.row_index) for e in row] == [(8.0, 0, 0), (2.0, 2, 0)]
    assert m.get_row_compact(2) == []


def test_matdim():
    d = MatDim.of((3, 4))
    assert d.rows == 3 and d.cols == 4 and d.transpose() == MatDim(4, 3) and tuple(d) == (3, 4)
    assert str(d) == "(rows: 3, cols: 4)"


def test_dimension_error_is_raised_before_any_device_work():
    m = Csr.from_data([[1, 2, 3]])
    with pytest.raises(MatError) as e:
        m.mul_dense(Dense.from_data([[1, 2]]))                              # sparse.rs:427-429
    assert e.value.kind == MatErr.IncorrectDimensions
    out = np.zeros(5)
    with pytest.raises(MatError) as e:
        Csr.from_data([[0, 0, 0, 0]] * 3).mul_vector(np.arange(5.0), out)   # sparse.rs:469-471
    assert e.value.kind == MatErr.IncorrectDimensions


def test_dense_static_kats():
    """The reference's own DenseS tests (dense_static.rs:74-96) and its from_data quirk (dims from the slices)."""
    a = DenseS.new_default(7, 5, np.int32)                                  # DenseS::<i32,7,5>::new_default()
    assert a == DenseS.from_data([[0] * 7] * 5, dtype=np.int32)
    assert a.get_dims() == MatDim(rows=7, cols=5)
    b = DenseS.from_data([[1, 2, 3], [4, 5, 6], [7, 8, 9]], 3, 3, np.int32)
    assert b.get_col(2).tolist() == [7, 8, 9]
    assert str(b) == "|    1    4    7|\n|    2    5    8|\n|    3    6    9|\n"
    assert DenseS.new(2.5, 2, 3).data.tolist() == [[2.5, 2.5]] * 3
    q = DenseS.from_data([[1, 2, 3], [4, 5, 6], [7, 8, 9]], 2, 2)           # window 2 x 2, dims 3 x 3 (dense_static.rs:22-33)
    assert q.data.tolist() == [[1, 2], [4, 5]] and q.get_dims() == MatDim(rows=3, cols=3)
    with pytest.raises(IndexError):
        DenseS.from_data([[1, 2]], 3, 1)                                     # slice shorter than ROWS: the reference's panic
    with pytest.raises(IndexError):
        q.get_col(2)


def test_mul_dense_s_front_end(golden, monkeypatch):
    """Csr::mul_dense_s (sparse.rs:448-466): dimension check against ROWS before any device work, then the same call as
    mul_dense with the DenseS columns.  The device call is replaced by the oracle HERE only to check which operand the
    front-end hands over (the GPU KAT is tests/test_gpu_parity.py::test_mul_dense_s_kat)."""
    from oracle.ref_cpu import OracleCsr
    k = golden["mul_dense"][0]
    m = Csr.from_data(k["csr_rows"])
    with pytest.raises(MatError) as e:
        m.mul_dense_s(DenseS.new_default(3, 3))                              # 4 columns against ROWS = 3
    assert e.value.kind == MatErr.IncorrectDimensions
    seen = {}

    def fake_mul_dense(self, rhs, algo="auto"):
        seen["dims"], seen["algo"] = rhs.get_dims(), algo
        o = OracleCsr.from_data(np.array(k["csr_rows"], np.float64), np.float64).mul_dense([np.ascontiguousarray(c, np.float64) for c in rhs.data])
        return Csr.from_raw_parts((self.dims.rows, rhs.col_count), o.v, o.col_index, o.row_index)

    monkeypatch.setattr(Csr, "mul_dense", fake_mul_dense)
    out = m.mul_dense_s(DenseS.from_data(k["dense_columns"], 4, 3), algo="vector")
    assert out == Csr.from_data(k["output_rows"]) and seen == {"dims": MatDim(rows=4, cols=3), "algo": "vector"}
    with pytest.raises(IndexError):                                          # col_count 3 > COLS 2: the reference's index panic
        m.mul_dense_s(DenseS.from_data(k["dense_columns"], 4, 2))


def test_integer_dtype_is_rejected_not_emulated():
    m = Csr.from_data([[1, 2]], dtype=np.int32)
    with pytest.raises(TypeError):
        m.mul_dense(Dense.from_data([[1, 2]], dtype=np.int32))


def test_no_cpu_fallback_without_device():
    """Without a usable GPU the product path must fail loudly."""
    from basic_sparse_matrix_b200 import gpu
    if gpu.device_count() > 0:
        pytest.skip("a CUDA device is present; the fallback check only applies to CPU-only hosts")
    m = Csr.from_data([[1.0, 2.0]])
    with pytest.raises(_lib.BsmError) as e:
        m.mul_dense(Dense.from_data([[1.0, 2.0]]))
    assert e.value.status == _lib.BSM_ERR_NO_DEVICE
