"""GPU (-m gpu): parity of the sm_100a kernels against the oracle, through the C ABI.

Bars (north_star / SURVEY §7.3-5):
  * vector-CSR kernel: BIT-EXACT against the restated sequential sum for ANY f32/f64 input
    (stored order, separately rounded multiply and add);
  * merge-path kernel: bit-exact on exactly representable (dyadic) inputs; otherwise
    |gpu - ref| <= tol * sum_j |a_ij b_jk| with tol = 1e-12 (f64) / 1e-5 (f32);
  * result Csr (zero-drop + finalise) identical field by field to the oracle's.
"""
import numpy as np
import pytest

from helpers import assert_bitwise, assert_tolerance, random_csr, random_dense
from basic_sparse_matrix_b200 import Csr, Dense, DenseS, MatErr, MatError, _lib, gen
from oracle import ref_numpy
from oracle.ref_cpu import OracleCsr

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}
DTYPES = [np.float64, np.float32]
NS = [1, 2, 3, 10, 31, 32, 33, 64, 128, 130, 256, 300]


def host_csr(dims, v, ci, ri):
    return Csr.from_raw_parts(dims, v, ci, ri)


def gpu_product(gpu, dims, v, ci, ri, b, algo, **tune):
    a = gpu.DeviceCsr.from_host(host_csr(dims, v, ci, ri))
    bd = gpu.DeviceDense.from_rowmajor(b)
    t = gpu.make_tuning(algo, **tune) if tune else None
    c = a.mul_dense(bd, algo=algo, tuning=t)
    out = c.to_rowmajor()
    info = gpu.last_launch_info()
    for h in (a, bd, c):
        h.close()
    return out, info


# ---- the reference's own tests, through the GPU path --------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_kats(gpu, golden, dtype):
    """test_dense_mul (sparse.rs:1082-1109) and test_nnz (1153-1178) read as in the reference:
    build operands with from_data, multiply, assert_eq! against from_data(expected)."""
    for k in golden["mul_dense"]:
        d = Dense.from_data(k["dense_columns"], dtype)
        s = Csr.from_data(k["csr_rows"], dtype)
        output_ref = Csr.from_data(k["output_rows"], dtype)
        output = s.mul_dense(d)
        assert output == output_ref, k["name"]
        if "nnz" in k:
            assert output.get_nnz() == k["nnz"]
        for algo in ("vector", "merge"):
            assert s.mul_dense(d, algo=algo) == output_ref, (k["name"], algo)


@pytest.mark.parametrize("dtype", DTYPES)
def test_mul_dense_s_kat(gpu, golden, dtype):
    """Csr::mul_dense_s (sparse.rs:448-466) on the test_dense_mul operands held as a DenseS<T,4,3>, as
    tests/cpp/test_reference_kats.cpp does through the C++ mirror."""
    k = golden["mul_dense"][0]
    m = Csr.from_data(k["csr_rows"], dtype)
    assert m.mul_dense_s(DenseS.from_data(k["dense_columns"], 4, 3, dtype)) == Csr.from_data(k["output_rows"], dtype)
    with pytest.raises(MatError) as e:
        m.mul_dense_s(DenseS.new_default(3, 3, dtype))
    assert e.value.kind == MatErr.IncorrectDimensions


@pytest.mark.parametrize("dtype", DTYPES)
def test_mul_vector_kat(gpu, golden, dtype):
    k = golden["mul_vector"]                                               # sparse.rs:1501-1529
    v = np.array(k["v"], dtype)
    out = np.zeros(5, dtype)
    with pytest.raises(MatError) as e:
        Csr.from_data(k["bad_dims_matrix"], dtype).mul_vector(v, out)
    assert e.value.kind == MatErr.IncorrectDimensions
    Csr.from_data(k["identity"], dtype).mul_vector(v, out)
    assert out.tolist() == k["v"]
    out = np.zeros(2, dtype)
    Csr.from_data(k["matrix"], dtype).mul_vector(v, out)
    assert out.tolist() == k["expected"]


def test_error_codes(gpu):
    a = gpu.DeviceCsr.from_host(Csr.from_data([[1.0, 2.0, 3.0]]))
    b = gpu.DeviceDense.from_rowmajor(np.ones((2, 4)))
    with pytest.raises(MatError) as e:
        a.mul_dense(b)
    assert e.value.kind == MatErr.IncorrectDimensions
    c = gpu.DeviceDense.alloc(5, 4)
    b3 = gpu.DeviceDense.from_rowmajor(np.ones((3, 4)))
    with pytest.raises(MatError):
        a.mul_dense(b3, out=c)                                             # C has the wrong shape
    with pytest.raises(_lib.BsmError):
        a.mul_dense(gpu.DeviceDense.from_rowmajor(np.ones((3, 4), np.float32)))   # dtype mismatch
    # not finalised -> MatrixNotFinalised; column out of range -> OutOfBounds
    m = Csr.new((2, 2))
    m.insert(1.0, 0, 0)
    with pytest.raises(MatError) as e:
        gpu.DeviceCsr.from_host(m)
    assert e.value.kind == MatErr.MatrixNotFinalised
    bad = Csr.from_raw_parts((2, 2), np.ones(1), np.array([5], np.uint64), np.array([0, 1, 1], np.uint64))
    with pytest.raises(MatError) as e:
        gpu.DeviceCsr.from_host(bad)
    assert e.value.kind == MatErr.OutOfBounds


# ---- vector kernel: bit-exact for any input -----------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", NS)
def test_vector_bitwise_random(gpu, dtype, n):
    rng = np.random.default_rng(100 + n)
    m, k = 517, 300
    v, ci, ri = random_csr(rng, m, k, dtype, mean_len=7, giant_row=40, giant_len=700)
    b = random_dense(rng, k, n, dtype)
    got, info = gpu_product(gpu, (m, k), v, ci, ri, b, "vector")
    assert info["algo"] == _lib.ALGO_VECTOR
    assert_bitwise(got, ref_numpy.mul_dense_rowmajor(v, ci, ri, b), f"vector n={n}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_vector_tuning_variants_are_bitwise_identical(gpu, dtype):
    rng = np.random.default_rng(5)
    m, k, n = 2049, 1000, 64
    v, ci, ri = random_csr(rng, m, k, dtype, mean_len=9)
    b = random_dense(rng, k, n, dtype)
    want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
    for tune in (dict(rows_per_slice=8), dict(rows_per_slice=64, stages=2), dict(rows_per_slice=256, stages=1),
                 dict(col_tile=16), dict(col_tile=32, ctas_per_sm=1), dict(prefer_wide_rows=1),
                 dict(warps_per_cta=2), dict(warps_per_cta=16, ctas_per_sm=1), dict(flags=_lib.TUNE_LITERAL),
                 dict(rows_per_warp=64, rows_per_slice=16), dict(rows_per_warp=96, rows_per_slice=8, reg_flavour=2),
                 dict(rows_per_warp=32, rows_per_slice=32, warps_per_cta=4, reg_flavour=3),
                 dict(rows_per_warp=512, rows_per_slice=4, stages=4, reg_flavour=4), dict(reg_flavour=1, stages=8),
                 # rows per warp that are a multiple of neither the slice nor 4 (short last slice, realigned row_ptr windows)
                 dict(rows_per_warp=50, rows_per_slice=16), dict(rows_per_warp=33, rows_per_slice=8, reg_flavour=5),
                 dict(rows_per_warp=251, rows_per_slice=32, lanes_per_row=32), dict(rows_per_warp=17, rows_per_slice=16, lanes_per_row=16)):
        got, _ = gpu_product(gpu, (m, k), v, ci, ri, b, "vector", **tune)
        assert_bitwise(got, want, f"tuning {tune}")


def _near_diagonal_csr(rng, m, k, dtype):
    """Rows mixing the tridiagonal part (columns row-1, row, row+1 — each present with probability 0.8,
    so every hit / miss / re-install sequence of the neighbour-row reuse occurs), far columns on both
    sides, repeated columns, unsorted rows and empty rows."""
    vals, cols, ri = [], [], [0]
    for r in range(m):
        kind = rng.integers(0, 10)
        row = []
        if kind > 0:                                            # kind 0: empty row
            row += [c for c in (r - 1, r, r + 1) if 0 <= c < k and rng.random() < 0.8]
            row += rng.integers(0, k, size=rng.integers(0, 5)).tolist()
            if kind == 1 and row:
                row += [row[0], row[-1]]                        # repeated columns
            row = sorted(row) if kind < 8 else list(rng.permutation(row))
        cols += row
        vals += rng.standard_normal(len(row)).tolist()
        ri.append(len(cols))
    return np.array(vals, dtype), np.array(cols, np.uint64), np.array(ri, np.uint64)


@pytest.mark.parametrize("dtype", DTYPES)
def test_vector_near_diagonal_rows_all_flavours_bitwise(gpu, dtype):
    """Stencil-like structure with every irregularity mixed in (missing neighbours, repeated and unsorted
    columns, empty rows), through every register flavour of the full-width shapes."""
    rng = np.random.default_rng(77)
    m = k = 3001
    per16 = 16 // np.dtype(dtype).itemsize
    v, ci, ri = _near_diagonal_csr(rng, m, k, dtype)
    for n in (32 * per16, 64 * per16, 32 * per16 // 2, 128 * per16):   # one, two tiles per lane, narrower vectors, four tiles
        b = random_dense(rng, k, n, dtype)
        want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
        for tune in (dict(), dict(reg_flavour=3), dict(reg_flavour=5), dict(reg_flavour=6), dict(reg_flavour=7, rows_per_slice=4, rows_per_warp=8),
                     dict(reg_flavour=8), dict(reg_flavour=8, rows_per_slice=64, stages=2, warps_per_cta=3)):
            got, info = gpu_product(gpu, (m, k), v, ci, ri, b, "vector", **tune)
            assert_bitwise(got, want, f"near-diagonal n={n} {tune} launched {info['reg_flavour']}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_vector_grouped_lanes_bitwise(gpu, dtype):
    """lanes_per_row < 32 with 2 or 4 register tiles per lane: 32/G lane groups each walk their own flat
    entry stream over a run of consecutive rows of the slice (rows of very different lengths side by side,
    empty groups, slices shorter than the group count)."""
    rng = np.random.default_rng(91)
    per16 = 16 // np.dtype(dtype).itemsize
    for m, k in ((3001, 3001), (5, 40), (1, 9), (67, 500)):
        mats = [_near_diagonal_csr(rng, m, k, dtype), random_csr(rng, m, k, dtype, mean_len=6, giant_row=min(3, m - 1), giant_len=200)]
        for mi, (v, ci, ri) in enumerate(mats):
            for g, nt in ((16, 2), (16, 4), (8, 2), (8, 4), (4, 2), (4, 4)):
                n = per16 * g * nt
                b = random_dense(rng, k, n, dtype)
                want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
                for tune in (dict(), dict(reg_flavour=7), dict(rows_per_slice=8, rows_per_warp=24), dict(rows_per_slice=64, stages=2, warps_per_cta=3)):
                    got, info = gpu_product(gpu, (m, k), v, ci, ri, b, "vector", lanes_per_row=g, **tune)
                    if mi == 0:   # (a giant row that cannot be staged falls back to a warp per row — by design)
                        assert (info["lanes_per_row"], info["reg_tiles"]) == (g, nt), info
                    assert_bitwise(got, want, f"grouped m={m} n={n} G={g} NT={nt} {tune}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_vector_flat_narrow_streams_bitwise(gpu, dtype):
    """64-byte output rows (four 128-bit lanes per row) as flat entry streams per lane group (reg_flavour 9, the default on short
    regular rows): rows of very different lengths side by side, empty rows and groups, slices shorter than the group count,
    rows per warp that are no multiple of the slice, a partial last column group (n = 6 of 8 f64 / 12 of 16 f32)."""
    rng = np.random.default_rng(123)
    per16 = 16 // np.dtype(dtype).itemsize
    for m, k in ((3001, 3001), (5, 40), (1, 9), (67, 500), (4099, 2000)):
        mats = [_near_diagonal_csr(rng, m, k, dtype), random_csr(rng, m, k, dtype, mean_len=6, giant_row=min(3, m - 1), giant_len=150)]
        for mi, (v, ci, ri) in enumerate(mats):
            for n in (4 * per16, 3 * per16):
                b = random_dense(rng, k, n, dtype)
                want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
                for tune in (dict(), dict(reg_flavour=9), dict(reg_flavour=9, rows_per_slice=8, rows_per_warp=24),
                             dict(reg_flavour=9, rows_per_slice=16, rows_per_warp=50), dict(reg_flavour=9, rows_per_slice=128, stages=2, warps_per_cta=3)):
                    got, info = gpu_product(gpu, (m, k), v, ci, ri, b, "vector", **tune)
                    if tune and mi == 0:   # (a giant row that cannot be staged falls back to the unstaged row-by-row kernel — by design)
                        assert (info["lanes_per_row"], info["reg_tiles"], info["reg_flavour"]) == (4, 1, 9), info
                    assert_bitwise(got, want, f"flat narrow m={m} n={n} {tune} launched {info}")
    # the default picks it for a stencil matrix
    a = gpu.DeviceCsr.laplacian(24, 24, 24, dtype=dtype)
    v, ci, ri, _ = gen.laplacian(24, 24, 24, dtype=dtype)
    n = 4 * per16
    b = gpu.DeviceDense.generate(24 ** 3, n, seed=5, mode=gen.MODE_REAL, dtype=dtype)
    c = a.mul_dense(b, algo="vector")
    assert gpu.last_launch_info()["reg_flavour"] == 9, gpu.last_launch_info()
    assert_bitwise(c.to_rowmajor(), ref_numpy.mul_dense_rowmajor(v, ci, ri, gen.dense_rows(24 ** 3, n, 5, gen.MODE_REAL, dtype=dtype)), "flat narrow default")
    for h in (a, b, c):
        h.close()


@pytest.mark.parametrize("g", [50, 37])
def test_vector_stencil_line_length_not_a_multiple_of_four(gpu, g):
    """3-D Laplacian on a g^3 grid with g = 50 / 37: the rows per warp follow the line length (50, 37), so warps start
    on rows that are not multiples of 4 and the last slice of a line is short. Bit-exact against the oracle."""
    a = gpu.DeviceCsr.laplacian(g, g, g)
    v, ci, ri, _ = gen.laplacian(g, g, g)
    for n in (64, 128, 32):
        b = gpu.DeviceDense.generate(g ** 3, n, seed=5, mode=gen.MODE_REAL)
        c = a.mul_dense(b, algo="vector")
        info = gpu.last_launch_info()
        want = ref_numpy.mul_dense_rowmajor(v, ci, ri, gen.dense_rows(g ** 3, n, 5, gen.MODE_REAL))
        assert_bitwise(c.to_rowmajor(), want, f"laplacian {g}^3 n={n} launched {info}")
        for h in (b, c):
            h.close()
    a.close()


def test_vector_slow_path_rows_longer_than_a_stage(gpu):
    rng = np.random.default_rng(9)
    m, k, n = 64, 5000, 32
    v, ci, ri = random_csr(rng, m, k, np.float64, mean_len=3, giant_row=17, giant_len=40_000)
    b = random_dense(rng, k, n, np.float64)
    got, _ = gpu_product(gpu, (m, k), v, ci, ri, b, "vector")
    assert_bitwise(got, ref_numpy.mul_dense_rowmajor(v, ci, ri, b), "slow path")


def test_line_length_is_a_majority_vote_over_all_rows(gpu):
    """The stencil line length comes from ALL rows (one statistics kernel, warp-aggregated votes), not from three sampled rows:
    a Laplacian whose rows at 1/4, 1/2 and 3/4 (round 1's sample points) are anything but stencil rows is still recognised;
    matrices without a majority are not; the column range and the longest row come from the same kernel."""
    g = 48
    v, ci, ri, dims = gen.laplacian(g, g, g)
    n = g ** 3
    with gpu.DeviceCsr.from_host(host_csr(dims, v, ci, ri)) as a:
        assert a.stats() == {"max_row_nnz": 7, "col_min": 0, "col_max": n - 1, "line_length": g}
    ci2 = ci.copy()
    for q in (1, 2, 3):                              # scramble the three rows round 1 sampled
        r = n // 4 * q
        s, e = int(ri[r]), int(ri[r + 1])
        ci2[s:e] = np.sort(np.random.default_rng(q).choice(n, e - s, replace=False)).astype(np.uint64)
    with gpu.DeviceCsr.from_host(host_csr(dims, v, ci2, ri)) as a:
        assert a.stats()["line_length"] == g
    # permuted copies of two stencils (line lengths 48 and 36 in equal parts): no majority -> no line length
    v2, c2, r2, d2 = gen.laplacian(36, 64, 48)      # same number of rows: 36*64*48 == 48^3
    assert d2[0] == n
    vv = np.concatenate([v[: int(ri[n // 2])], v2[int(r2[n // 2]):]])
    cc = np.concatenate([ci[: int(ri[n // 2])], c2[int(r2[n // 2]):]])
    rr = np.concatenate([ri[: n // 2], r2[n // 2:] - r2[n // 2] + ri[n // 2]])
    with gpu.DeviceCsr.from_host(host_csr(dims, vv, cc, rr)) as a:
        assert a.stats()["line_length"] in (0, g, 36)   # exactly half each: whichever has >= half of the rows, or none
    rng = np.random.default_rng(5)
    vr, cr, rrr = random_csr(rng, 9000, 7000, np.float64, mean_len=6)
    with gpu.DeviceCsr.from_host(host_csr((9000, 7000), vr, cr, rrr)) as a:
        st = a.stats()
        assert st["line_length"] == 0 and st["col_max"] == int(cr.max()) and st["col_min"] == int(cr.min())
        assert st["max_row_nnz"] == int(np.diff(rrr.astype(np.int64)).max())


# ---- row-block kernel (band-like matrices) -----------------------------------------------------------------
def _runs_csr(rng, m, k, dtype, max_len, empty_frac=0.1, band=True):
    """Every row stores one run of consecutive columns (a band when band=True, else runs anywhere)."""
    vals, cols, ri = [], [], [0]
    for r in range(m):
        if rng.random() >= empty_frac:
            ln = int(rng.integers(1, max_len + 1))
            lo = max(0, min(k - ln, (r * k // m) - ln // 2)) if band else int(rng.integers(0, k - ln + 1))
            ln = min(ln, k - lo)
            cols += list(range(lo, lo + ln))
            vals += rng.standard_normal(ln).tolist()
        ri.append(len(cols))
    return np.array(vals, dtype), np.array(cols, np.uint64), np.array(ri, np.uint64)


@pytest.mark.parametrize("dtype", DTYPES)
def test_rowblock_bitwise(gpu, dtype):
    """BSM_ALGO_ROWBLOCK: blocks of 8 consecutive rows share their B-row loads; every row still sums in stored
    order with the unfused multiply-add -> bit-exact for any values (ragged bands, empty rows, runs anywhere)."""
    rng = np.random.default_rng(606)
    for m, k, max_len, band in ((1003, 1003, 40, True), (77, 300, 65, True), (5, 9, 9, True), (1, 4, 3, True), (400, 5000, 30, False)):
        v, ci, ri = _runs_csr(rng, m, k, dtype, max_len, band=band)
        for n in (1, 3, 8, 32, 33, 64, 100, 130):
            b = random_dense(rng, k, n, dtype)
            got, info = gpu_product(gpu, (m, k), v, ci, ri, b, "rowblock")
            assert info["algo"] == _lib.ALGO_ROWBLOCK
            assert_bitwise(got, ref_numpy.mul_dense_rowmajor(v, ci, ri, b), f"rowblock m={m} n={n} band={band}")
        got, _ = gpu_product(gpu, (m, k), v, ci, ri, random_dense(np.random.default_rng(1), k, 64, dtype), "rowblock", col_tile=16)
        assert_bitwise(got, ref_numpy.mul_dense_rowmajor(v, ci, ri, random_dense(np.random.default_rng(1), k, 64, dtype)), "rowblock col_tile")
        # full-width 128-bit shapes: two register tiles per lane on half the lanes by default, one tile per lane on request
        per16 = 16 // np.dtype(dtype).itemsize
        for lanes in (8, 16, 32):
            n = per16 * lanes
            b = random_dense(rng, k, n, dtype)
            want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
            got, info = gpu_product(gpu, (m, k), v, ci, ri, b, "rowblock")
            assert (info["lanes_per_row"], info["reg_tiles"]) == (lanes // 2, 2), info
            assert_bitwise(got, want, f"rowblock two tiles m={m} n={n}")
            for rb in (4, 8):
                got, info = gpu_product(gpu, (m, k), v, ci, ri, b, "rowblock", lanes_per_row=lanes, rows_per_slice=rb)
                assert (info["lanes_per_row"], info["reg_tiles"]) == (lanes, 1), info
                assert_bitwise(got, want, f"rowblock one tile m={m} n={n} rb={rb}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_rowblock_fused_is_within_tolerance(gpu, dtype):
    """BSM_TUNE_FUSED (opt-in, never a default): the row-block kernel with one FMA per product instead of the reference's
    separately rounded multiply and add. Agreement is then the stated tolerance (1e-12 f64 / 1e-5 f32 of sum|a*b|), not bitwise;
    on exactly representable data it is still bit-exact. Shapes without a fused variant are refused, not silently unfused."""
    rng = np.random.default_rng(808)
    m = k = 2003
    v, ci, ri = _runs_csr(rng, m, k, dtype, 65, band=True)
    fused = _lib.TUNE_A_EVICT_FIRST | _lib.TUNE_C_STREAMING | _lib.TUNE_FUSED
    per16 = 16 // np.dtype(dtype).itemsize
    for lanes in (8, 32):
        n = per16 * lanes
        b = random_dense(rng, k, n, dtype)
        for tune in (dict(rows_per_slice=4, lanes_per_row=lanes), dict(rows_per_slice=8, lanes_per_row=lanes), dict()):   # one tile per lane (4 / 8 rows), two tiles
            got, info = gpu_product(gpu, (m, k), v, ci, ri, b, "rowblock", flags=fused, **tune)
            assert info["algo"] == _lib.ALGO_ROWBLOCK and info["reg_tiles"] == (1 if tune else 2)
            assert_tolerance(got, ref_numpy.mul_dense_rowmajor(v, ci, ri, b), ref_numpy.abs_product_sum(v, ci, ri, b), TOL[dtype], f"fused n={n} {tune}")
        ve = (np.round(v * 8) / 8).astype(dtype)
        be = random_dense(rng, k, n, dtype, exact=True)
        got, _ = gpu_product(gpu, (m, k), ve, ci, ri, be, "rowblock", flags=fused)
        assert_bitwise(got, ref_numpy.mul_dense_rowmajor(ve, ci, ri, be), f"fused, exact data n={n}")
    with pytest.raises(_lib.BsmError) as e:       # 3 columns: no 128-bit lanes, no fused variant
        gpu_product(gpu, (m, k), v, ci, ri, random_dense(rng, k, 3, dtype), "rowblock", flags=fused)
    assert e.value.status == _lib.BSM_ERR_NOT_SUPPORTED


def test_rowblock_refuses_rows_that_are_not_runs_and_auto_picks_it_for_bands(gpu):
    rng = np.random.default_rng(607)
    v, ci, ri = random_csr(rng, 300, 300, np.float64, mean_len=6)
    with pytest.raises(_lib.BsmError):
        gpu_product(gpu, (300, 300), v, ci, ri, random_dense(rng, 300, 8, np.float64), "rowblock")
    # a band large enough for the heuristic: AUTO runs the row-block kernel, bit-identical to the vector kernel
    a = gpu.DeviceCsr.band(40_000, 16, dtype=np.float32)
    bd = gpu.DeviceDense.generate(40_000, 32, seed=9, mode=gen.MODE_REAL, dtype=np.float32)
    auto = a.mul_dense(bd)
    assert gpu.last_launch_info()["algo"] == _lib.ALGO_ROWBLOCK
    vec = a.mul_dense(bd, algo="vector")
    assert gpu.last_launch_info()["algo"] == _lib.ALGO_VECTOR
    assert_bitwise(auto.to_rowmajor(), vec.to_rowmajor(), "rowblock vs vector on a band")
    for h in (a, bd, auto, vec):
        h.close()


# ---- merge-path kernel -------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", NS)
def test_merge_bitwise_on_exact_inputs(gpu, dtype, n):
    rng = np.random.default_rng(200 + n)
    m, k = 700, 256
    v, ci, ri = random_csr(rng, m, k, dtype, mean_len=5, empty_frac=0.3, giant_row=333, giant_len=3000, exact=True)
    b = random_dense(rng, k, n, dtype, exact=True)
    got, info = gpu_product(gpu, (m, k), v, ci, ri, b, "merge")
    assert info["algo"] == _lib.ALGO_MERGE
    assert_bitwise(got, ref_numpy.mul_dense_rowmajor(v, ci, ri, b), f"merge exact n={n}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 10, 64, 128])
def test_merge_within_tolerance_on_real_inputs(gpu, dtype, n):
    rng = np.random.default_rng(300 + n)
    m, k = 900, 400
    v, ci, ri = random_csr(rng, m, k, dtype, mean_len=6, giant_row=1, giant_len=5000)
    b = random_dense(rng, k, n, dtype)
    got, _ = gpu_product(gpu, (m, k), v, ci, ri, b, "merge")
    want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
    scale = ref_numpy.abs_product_sum(v, ci, ri, b)
    assert_tolerance(got, want, scale, TOL[dtype], f"merge real n={n}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_merge_grouped_lanes(gpu, dtype):
    """merge-path with lanes_per_row < 32 and 2 register tiles per lane (32/G chunks side by side per warp):
    bit-exact on exactly representable data, within tolerance on real data. (The 4-tile and 4-lane variants of round 1
    never won a sweep and are no longer built.)"""
    rng = np.random.default_rng(404)
    m, k = 1500, 600
    per16 = 16 // np.dtype(dtype).itemsize
    ve, cie, rie = random_csr(rng, m, k, dtype, mean_len=4, empty_frac=0.4, giant_row=777, giant_len=9000, exact=True)
    vr, cir, rir = random_csr(rng, m, k, dtype, mean_len=6, giant_row=1, giant_len=5000)
    for g, nt in ((16, 2), (8, 2)):
        n = per16 * g * nt
        be = random_dense(rng, k, n, dtype, exact=True)
        for tune in (dict(), dict(merge_items=96), dict(merge_items=32, warps_per_cta=4)):
            got, info = gpu_product(gpu, (m, k), ve, cie, rie, be, "merge", lanes_per_row=g, **tune)
            assert (info["algo"], info["lanes_per_row"], info["reg_tiles"]) == (_lib.ALGO_MERGE, g, nt), info
            assert_bitwise(got, ref_numpy.mul_dense_rowmajor(ve, cie, rie, be), f"merge grouped exact G={g} NT={nt} {tune}")
        br = random_dense(rng, k, n, dtype)
        got, _ = gpu_product(gpu, (m, k), vr, cir, rir, br, "merge", lanes_per_row=g)
        assert_tolerance(got, ref_numpy.mul_dense_rowmajor(vr, cir, rir, br), ref_numpy.abs_product_sum(vr, cir, rir, br), TOL[dtype],
                         f"merge grouped real G={g} NT={nt}")


def test_merge_items_variants(gpu):
    rng = np.random.default_rng(17)
    m, k, n = 1500, 600, 64
    v, ci, ri = random_csr(rng, m, k, np.float64, mean_len=4, empty_frac=0.5, giant_row=1499, giant_len=9000, exact=True)
    b = random_dense(rng, k, n, np.float64, exact=True)
    want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
    for tune in (dict(merge_items=16), dict(merge_items=64), dict(merge_items=512, warps_per_cta=4),
                 dict(col_tile=16), dict(prefer_wide_rows=-1)):
        got, _ = gpu_product(gpu, (m, k), v, ci, ri, b, "merge", **tune)
        assert_bitwise(got, want, f"merge tuning {tune}")


def test_merge_is_deterministic(gpu):
    """A hub row of 20 000 entries with the DEFAULT items per chunk: more than 64 carrying chunks, so the long-run fix-up
    kernel (merge_fixup_long_kernel) sums it. Run-to-run identical, and against the oracle: within the stated tolerance on
    real data, bit-exact on dyadic data."""
    rng = np.random.default_rng(23)
    m, k, n = 800, 500, 64
    v, ci, ri = random_csr(rng, m, k, np.float64, giant_row=10, giant_len=20_000)
    b = random_dense(rng, k, n, np.float64)
    first, info = gpu_product(gpu, (m, k), v, ci, ri, b, "merge")
    assert info["kernels"] >= 3 and (20_000 + 11) // info["merge_items"] > 64, info   # merge + both fix-up kernels
    for _ in range(3):
        again, _ = gpu_product(gpu, (m, k), v, ci, ri, b, "merge")
        assert_bitwise(again, first, "run-to-run")
    assert_tolerance(first, ref_numpy.mul_dense_rowmajor(v, ci, ri, b), ref_numpy.abs_product_sum(v, ci, ri, b), 1e-12, "long-run fix-up vs oracle")
    ve, cie, rie = random_csr(rng, m, k, np.float64, giant_row=10, giant_len=20_000, exact=True)
    be = random_dense(rng, k, n, np.float64, exact=True)
    got, _ = gpu_product(gpu, (m, k), ve, cie, rie, be, "merge")
    assert_bitwise(got, ref_numpy.mul_dense_rowmajor(ve, cie, rie, be), "long-run fix-up, exact data")


# ---- edge shapes ----------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["vector", "merge", "auto"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_edge_shapes(gpu, algo, dtype):
    cases = []
    cases.append(((1, 1), np.array([2.0]), [0], [0, 1]))                          # 1x1
    cases.append(((4, 3), np.zeros(0), [], [0, 0, 0, 0, 0]))                      # no entries at all
    cases.append(((5, 4), np.array([1.0, 2.0, 3.0]), [3, 0, 0], [0, 0, 0, 3, 3, 3]))   # empty rows top and bottom, dup col
    cases.append(((3, 6), np.arange(1.0, 13.0), [5, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5], [0, 6, 6, 12]))  # unsorted, empty middle
    cases.append(((2, 2), np.array([1.0, -1.0]), [0, 0], [0, 2, 2]))              # duplicates cancel to exact 0
    for dims, v, ci, ri in cases:
        for n in (1, 2, 5, 64):
            rng = np.random.default_rng(n)
            b = random_dense(rng, dims[1], n, dtype, exact=True)
            v_ = v.astype(dtype)
            got, _ = gpu_product(gpu, dims, v_, np.array(ci, np.uint64), np.array(ri, np.uint64), b, algo)
            want = ref_numpy.mul_dense_rowmajor(v_, np.array(ci, np.uint64), np.array(ri, np.uint64), b)
            assert_bitwise(got, want, f"{dims} n={n} {algo}")


def test_single_giant_row_auto_uses_merge(gpu):
    rng = np.random.default_rng(31)
    a, x = gen.bench_as_written(20_000)          # the reference bench shape: ~99 % of nnz in one row
    v, ci, ri = a.raw_parts()
    b = x.to_rowmajor()
    got, info = gpu_product(gpu, (1000, 1000), v, ci, ri, b, "auto")
    assert info["algo"] == _lib.ALGO_MERGE
    assert_bitwise(got, ref_numpy.mul_dense_rowmajor(v, ci, ri, b), "bench as written")   # integers: exact


# ---- result Csr: zero-drop + finalise ----------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_result_csr_matches_oracle_field_by_field(gpu, dtype):
    rng = np.random.default_rng(41)
    m, k, n = 120, 80, 7
    v, ci, ri = random_csr(rng, m, k, dtype, exact=True, empty_frac=0.4)
    b = random_dense(rng, k, n, dtype, exact=True)
    b[:, 2] = 0                                   # a whole zero output column
    out = host_csr((m, k), v, ci, ri).mul_dense(Dense.from_data([b[:, c] for c in range(n)], dtype))
    ref = OracleCsr.from_raw((m, k), v, ci, ri).mul_dense([b[:, c].copy() for c in range(n)])
    assert out.is_finalised and out.get_dims().rows == m and out.get_dims().cols == n
    assert_bitwise(out.v, ref.v)
    assert np.array_equal(out.col_index, ref.col_index) and np.array_equal(out.row_index, ref.row_index)
    assert out.get_nnz() == ref.get_nnz() < m * n


def test_zero_drop_keeps_nan_drops_negative_zero(gpu):
    d = np.array([[0.0, -0.0, np.nan, 1.0], [-0.0, 0.0, 0.0, 0.0], [2.0, 0.0, np.inf, -3.0]])
    dd = gpu.DeviceDense.from_rowmajor(d)
    r = dd.into_csr().to_host()
    rv, rc, rr = ref_numpy.dense_to_csr(d)
    assert r.row_index.tolist() == rr.tolist() == [0, 2, 2, 5]
    assert r.col_index.tolist() == rc.tolist() == [2, 3, 0, 2, 3]
    assert np.isnan(r.v[0]) and r.v[1:].tolist() == [1.0, 2.0, np.inf, -3.0]


def test_compaction_large(gpu):
    rng = np.random.default_rng(43)
    d = rng.integers(-1, 2, size=(70_001, 33)).astype(np.float32)    # ~1/3 zeros, crosses scan tiles
    r = gpu.DeviceDense.from_rowmajor(d).into_csr().to_host()
    rv, rc, rr = ref_numpy.dense_to_csr(d)
    assert np.array_equal(r.v, rv) and np.array_equal(r.col_index, rc) and np.array_equal(r.row_index, rr)


# ---- host <-> device format conversion -----------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_upload_download_roundtrip(gpu, dtype):
    rng = np.random.default_rng(47)
    v, ci, ri = random_csr(rng, 333, 77, dtype)
    back = gpu.DeviceCsr.from_host(host_csr((333, 77), v, ci, ri)).to_host()
    assert back == host_csr((333, 77), v, ci, ri)
    for rows, cols in ((1, 1), (100, 1), (65, 33), (1000, 10), (31, 130)):
        d = Dense.from_data([rng.standard_normal(rows).astype(dtype) for _ in range(cols)], dtype)
        dev = gpu.DeviceDense.from_host(d)
        assert dev.to_host() == d                                       # column-major round trip
        assert np.array_equal(dev.to_rowmajor(), d.to_rowmajor())       # device layout is row-major


def test_row_slice_upload(gpu):
    rng = np.random.default_rng(53)
    m, k, n = 400, 120, 16
    v, ci, ri = random_csr(rng, m, k, np.float64)
    b = random_dense(rng, k, n, np.float64)
    want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
    bounds = gpu.partition_rows(ri, 3).astype(int)
    bd = gpu.DeviceDense.from_rowmajor(b)
    parts = []
    for p in range(3):
        a = gpu.DeviceCsr.from_host(host_csr((m, k), v, ci, ri), int(bounds[p]), int(bounds[p + 1]))
        parts.append(a.mul_dense(bd, algo="vector").to_rowmajor())
    assert_bitwise(np.concatenate(parts, axis=0), want, "row-partitioned product")


# ---- device generators == numpy generators --------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_device_generators_match_numpy(gpu, dtype):
    for (nx, ny, nz, r0, r1) in ((9, 7, 1, 0, None), (6, 5, 4, 17, 101)):
        a = gpu.DeviceCsr.laplacian(nx, ny, nz, r0, r1, dtype).to_host()
        v, ci, ri, dims = gen.laplacian(nx, ny, nz, r0, r1, dtype)
        assert a == host_csr(dims, v, ci, ri)
    a = gpu.DeviceCsr.band(300, 32, 10, 290, dtype).to_host()
    v, ci, ri, dims = gen.band(300, 32, 10, 290, dtype)
    assert a == host_csr(dims, v, ci, ri)
    for mode in (gen.MODE_EXACT, gen.MODE_REAL, gen.MODE_EXACT_SMALL):
        a = gpu.DeviceCsr.rmat(10, 30_000, seed=3, mode=mode, dtype=dtype).to_host()
        v, ci, ri, dims = gen.rmat(10, 30_000, seed=3, mode=mode, dtype=dtype)
        assert a == host_csr(dims, v, ci, ri)
        d = gpu.DeviceDense.generate(123, 17, seed=4, mode=mode, offset=0.5, dtype=dtype).to_rowmajor()
        assert np.array_equal(d, gen.dense_rows(123, 17, 4, mode, 0.5, dtype))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 5, 32, 33, 100])
def test_pipelined_host_to_host_dense_product(gpu, dtype, n):
    """bsm_mul_dense_host_dense_*: the pipelined host-to-host product with a dense result == the oracle, bit for bit."""
    rng = np.random.default_rng(21)
    m, k = 777, 501
    v, ci, ri = random_csr(rng, m, k, dtype, mean_len=7)
    b = random_dense(rng, k, n, dtype)
    a = host_csr((m, k), v, ci, ri)
    out = a.mul_dense_into(Dense.from_data([b[:, c] for c in range(n)], dtype))
    assert_bitwise(out.to_rowmajor(), ref_numpy.mul_dense_rowmajor(v, ci, ri, b), f"host pipeline n={n}")
    with pytest.raises(MatError) as e:
        a.mul_dense_into(Dense.new_default_with_dims(2, k + 1, dtype))
    assert e.value.kind == MatErr.IncorrectDimensions


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kind", ["random", "band", "hub"])
def test_pipelined_literal_call_many_blocks_and_chunks(gpu, dtype, kind, monkeypatch):
    """bsm_mul_dense_host_into_* and bsm_mul_dense_host_dense_* with the pipeline's block / chunk sizes shrunk so that a small
    product runs through dozens of row blocks and B chunks (views of A, windows of B, per-block compaction, the lagged
    device->host copies): the result Csr must equal the oracle's field by field, the dense twin bit for bit."""
    monkeypatch.setenv("BSM_PIPE_BLOCK_BYTES", "20000")
    monkeypatch.setenv("BSM_PIPE_CHUNK_BYTES", "9000")
    rng = np.random.default_rng(91)
    m, k, n = 1501, 1203, 12
    if kind == "band":
        v, ci, ri = _runs_csr(rng, m, k, dtype, 30, band=True)
        v = (np.round(v * 8) / 8).astype(dtype)
        v[v == 0] = 0.125
    elif kind == "hub":
        v, ci, ri = random_csr(rng, m, k, dtype, mean_len=5, empty_frac=0.3, giant_row=700, giant_len=6000, exact=True)
    else:
        v, ci, ri = random_csr(rng, m, k, dtype, mean_len=7, exact=True)
    b = random_dense(rng, k, n, dtype, exact=True)
    b[:, 3] = 0                                   # a zero output column: dropped from every row of the result Csr
    a = host_csr((m, k), v, ci, ri)
    rhs = Dense.from_data([b[:, c] for c in range(n)], dtype)
    ref = OracleCsr.from_raw((m, k), v, ci, ri).mul_dense([b[:, c].copy() for c in range(n)])
    out = a.mul_dense(rhs)                        # allocating form
    assert_bitwise(out.v, ref.v)
    assert np.array_equal(out.col_index, ref.col_index) and np.array_equal(out.row_index, ref.row_index)
    ov, oc, orow = np.empty(m * n, dtype), np.empty(m * n, np.uint64), np.empty(m + 1, np.uint64)
    out2 = a.mul_dense_csr_into(rhs, ov, oc, orow)
    assert out2 == out
    with pytest.raises(_lib.BsmError):            # result arrays too small
        a.mul_dense_csr_into(rhs, ov[:10], oc[:10], orow)
    dense = a.mul_dense_into(rhs)
    assert_bitwise(dense.to_rowmajor(), ref_numpy.mul_dense_rowmajor(v, ci, ri, b), f"dense twin {kind}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("threads", ["0", "3"])
def test_literal_call_columns_travel_as_row_masks(gpu, dtype, threads, monkeypatch):
    """The literal call sends one keep-bit per output instead of a usize column and host threads expand the masks into col_index
    (BSM_PIPE_EXPAND_THREADS; 0 = the device writes usize columns and they are copied). Full rows (one pattern copy), rows with a
    dropped zero, empty rows, many blocks: the result must equal the oracle's field by field, with and without the masks."""
    monkeypatch.setenv("BSM_PIPE_BLOCK_BYTES", "30000")
    monkeypatch.setenv("BSM_PIPE_CHUNK_BYTES", "9000")
    monkeypatch.setenv("BSM_PIPE_EXPAND_THREADS", threads)
    monkeypatch.setenv("BSM_PIPE_EXPAND_MIN_ENTRIES", "0")         # (small results skip the masks by default)
    rng = np.random.default_rng(5)
    m, k, n = 4001, 900, 20
    v, ci, ri = _runs_csr(rng, m, k, dtype, 12, empty_frac=0.0, band=True)   # every row has entries: full result rows ...
    v = (np.round(v * 8) / 8).astype(dtype)
    v[v == 0] = 0.125
    b = random_dense(rng, k, n, dtype, exact=True)
    b[b == 0] = 1.0
    b[600:640, 7] = 0                                               # ... except where a window of B is zero in one column
    drop = set(range(1000, 1100)) | {2500, 3999}                    # ... and in blocks with empty rows
    parts = [(v[int(ri[r]):int(ri[r + 1])], ci[int(ri[r]):int(ri[r + 1])]) if r not in drop else (v[:0], ci[:0]) for r in range(m)]
    v, ci = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    ri = np.concatenate([[0], np.cumsum([len(p[0]) for p in parts])]).astype(np.uint64)
    a = host_csr((m, k), v, ci, ri)
    rhs = Dense.from_data([b[:, c] for c in range(n)], dtype)
    ref = OracleCsr.from_raw((m, k), v, ci, ri).mul_dense([b[:, c].copy() for c in range(n)])
    full = np.diff(ref.row_index.astype(np.int64)) == n
    assert 0.5 < full.mean() < 1.0                                  # most rows full, some not
    ov, oc, orow = np.empty(m * n, dtype), np.full(m * n, 0xDEADBEEF, np.uint64), np.empty(m + 1, np.uint64)
    out = a.mul_dense_csr_into(rhs, ov, oc, orow)
    assert_bitwise(out.v, ref.v)
    assert np.array_equal(out.col_index, ref.col_index) and np.array_equal(out.row_index, ref.row_index)
    assert a.mul_dense(rhs) == out                                  # allocating form


def test_literal_call_edge_shapes(gpu):
    """Empty rows at both ends, one column, zero columns, 1 x 1."""
    for (m, k, n) in ((7, 5, 0), (9, 4, 1), (3, 3, 2), (1, 1, 1)):   # (a 0-row Csr cannot be finalised in the reference: "big eek")
        a = Csr.new((m, k), np.float64)
        if m >= 3:
            a.insert(2.0, 1, 0)
        a = a.finalise()
        rhs = Dense.from_data([np.arange(1, k + 1, dtype=np.float64) for _ in range(n)]) if n else Dense(0, k, [])
        out = a.mul_dense(rhs)
        want = Csr.new((m, n), np.float64)
        if m >= 3:
            for c in range(n):
                want.insert(2.0, 1, c)
        assert out == want.finalise(), (m, k, n)


@pytest.mark.parametrize("seed", range(6))
def test_random_shape_sweep(gpu, seed):
    """Randomised shapes, densities, dtypes and kernels (seeded): every draw against the oracle — bit-exact
    for the vector kernel, bit-exact on dyadic data / within tolerance on real data for merge-path, and
    identical results from `auto`."""
    rng = np.random.default_rng(1000 + seed)
    for _ in range(8):
        dtype = DTYPES[int(rng.integers(0, 2))]
        m = int(rng.choice([1, 2, 7, 33, 100, 513, 2000, 5001]))
        k = int(rng.choice([1, 3, 64, 1000, 4099]))
        n = int(rng.choice([1, 2, 5, 8, 16, 24, 32, 48, 64, 96, 128, 200]))
        mean_len = float(rng.choice([0.5, 2, 7, 30, 120]))
        giant = bool(rng.integers(0, 2)) and m > 2
        exact = bool(rng.integers(0, 2))
        v, ci, ri = random_csr(rng, m, k, dtype, mean_len=min(mean_len, 4.0 * k), exact=exact,
                               giant_row=int(rng.integers(0, m)) if giant else None, giant_len=int(rng.integers(50, 3000)) if giant else 0)
        b = random_dense(rng, k, n, dtype, exact=exact)
        want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
        tag = f"seed={seed} {np.dtype(dtype).name} m={m} k={k} n={n} mean={mean_len} giant={giant} exact={exact}"
        got_v, _ = gpu_product(gpu, (m, k), v, ci, ri, b, "vector")
        assert_bitwise(got_v, want, "vector " + tag)
        got_m, _ = gpu_product(gpu, (m, k), v, ci, ri, b, "merge")
        if exact:
            assert_bitwise(got_m, want, "merge " + tag)
        else:
            scale = ref_numpy.mul_dense_rowmajor(np.abs(v), ci, ri, np.abs(b))
            assert_tolerance(got_m, want, scale, TOL[dtype], "merge " + tag)
        got_a, info = gpu_product(gpu, (m, k), v, ci, ri, b, "auto")
        assert_bitwise(got_a, got_v if info["algo"] == _lib.ALGO_VECTOR else got_m, "auto " + tag)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 10, 32, 64, 128, 200])
def test_scatter_variant_writes_every_destination(gpu, dtype, n):
    """bsm_spmm_scatter (multiply + all-gather fused, P2P stores): on one GPU the "peers" are three local
    full-size buffers; each must receive this rank's rows at row_offset, bit-identical to the plain
    product, and nothing else."""
    rng = np.random.default_rng(33)
    m, k, off, total = 700, 900, 123, 1000
    v, ci, ri = random_csr(rng, m, k, dtype, mean_len=6)
    b = random_dense(rng, k, n, dtype)
    want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
    a = gpu.DeviceCsr.from_host(host_csr((m, k), v, ci, ri))
    bd = gpu.DeviceDense.from_rowmajor(b)
    sentinel = np.full((total, n), 7.25, dtype)
    fulls = [gpu.DeviceDense.from_rowmajor(sentinel) for _ in range(3)]
    a.mul_dense_scatter(bd, fulls, off, algo="vector")
    for f in fulls:
        got = f.to_rowmajor()
        assert_bitwise(got[off:off + m], want, f"scatter n={n}")
        assert np.all(got[:off] == 7.25) and np.all(got[off + m:] == 7.25)
    # dimension and kernel-family errors
    with pytest.raises(MatError) as e:
        a.mul_dense_scatter(bd, fulls, total - m + 1)
    assert e.value.kind == MatErr.IncorrectDimensions
    with pytest.raises(_lib.BsmError) as e2:
        a.mul_dense_scatter(bd, fulls, off, algo="merge")
    assert e2.value.status == _lib.BSM_ERR_NOT_SUPPORTED
    for h in fulls + [a, bd]:
        h.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("algo", ["vector", "merge"])
def test_borrowed_strided_operands(gpu, dtype, algo):
    """B and C as borrowed views with a leading dimension larger than the column count and an odd
    column offset (16-byte alignment broken): the lane shape must fall back to narrower loads and the
    product must still be bit-exact (dyadic data, so merge-path too); padding columns stay untouched."""
    rng = np.random.default_rng(44)
    m, k = 300, 257
    s = np.dtype(dtype).itemsize
    v, ci, ri = random_csr(rng, m, k, dtype, mean_len=5, exact=True)
    a = gpu.DeviceCsr.from_host(host_csr((m, k), v, ci, ri))
    for n, ld, off in ((64, 69, 3), (32, 40, 1), (7, 16, 0), (128, 128 + 2, 2)):
        b_full = random_dense(rng, k, ld, dtype, exact=True)
        c_full = np.full((m, ld), -3.5, dtype)
        bd, cd = gpu.DeviceDense.from_rowmajor(b_full), gpu.DeviceDense.from_rowmajor(c_full)
        bi, cinfo = bd.info(), cd.info()
        bv = gpu.DeviceDense.borrow(bi["ptr"] + off * s, k, n, bi["ld"], dtype)
        cv = gpu.DeviceDense.borrow(cinfo["ptr"] + off * s, m, n, cinfo["ld"], dtype)
        a.mul_dense(bv, out=cv, algo=algo)
        got = cd.to_rowmajor()
        want = ref_numpy.mul_dense_rowmajor(v, ci, ri, np.ascontiguousarray(b_full[:, off:off + n]))
        assert_bitwise(got[:, off:off + n], want, f"strided {algo} n={n} ld={ld} off={off}")
        assert np.all(got[:, :off] == -3.5) and np.all(got[:, off + n:] == -3.5)
        for h in (bv, cv, bd, cd):
            h.close()
    a.close()


# (last in the file: first exercised on a GPU by the round-end run)
@pytest.mark.parametrize("dtype", DTYPES)
def test_odd_wide_column_counts(gpu, dtype):
    """n = 129, 131, 255, 257: alignment forces one-element lanes, more of them than one pass holds; the pass is split
    (fit_pass_width) instead of dropping the columns past 32 lanes x 4 tiles. Every kernel family."""
    rng = np.random.default_rng(515)
    m, k = 700, 300
    v, ci, ri = random_csr(rng, m, k, dtype, mean_len=6, giant_row=5, giant_len=900, exact=True)
    vb, cb, rb = _runs_csr(rng, m, k, dtype, 40)
    for n in (129, 131, 255, 257):
        b = random_dense(rng, k, n, dtype, exact=True)
        want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
        for algo in ("vector", "merge", "auto"):
            got, info = gpu_product(gpu, (m, k), v, ci, ri, b, algo)
            assert info["passes"] >= 2
            assert_bitwise(got, want, f"n={n} {algo}")
        got, _ = gpu_product(gpu, (m, k), vb, cb, rb, b, "rowblock")
        assert_bitwise(got, ref_numpy.mul_dense_rowmajor(vb, cb, rb, b), f"n={n} rowblock")
