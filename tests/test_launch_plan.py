"""CPU: the launch heuristics of the vector kernel, through bsm_plan_vector (a dry run: no device needed).

Every expectation below is a configuration that was MEASURED on a B200 this round (profiles/README.md); the
tests pin the plan, so that a later change to the heuristics that silently reverts one of them fails here
rather than in a benchmark. B200: 148 SMs, 227 KB opt-in shared memory per CTA."""
import ctypes as C

import pytest

from basic_sparse_matrix_b200 import _lib

SMS, SMEM = 148, 232448
F32, F64 = _lib.dtype_code("float32"), _lib.dtype_code("float64")


def plan(dtype, rows, nnz, max_row, stride, n, **tune):
    t = _lib.Tuning()
    for k, v in tune.items():
        setattr(t, k, v)
    info = _lib.LaunchInfo()
    st = _lib.lib().bsm_plan_vector(dtype, rows, nnz, max_row, stride, n, C.byref(t), SMS, SMEM, C.byref(info))
    assert st == 0, _lib.lib().bsm_last_error_string()
    return info.as_dict()


L3D = dict(rows=256 ** 3, nnz=117_047_296, max_row=7)          # 3-D 7-point Laplacian 256^3


def test_headline_x128_f64():
    p = plan(F64, L3D["rows"], L3D["nnz"], L3D["max_row"], 256, 128)
    assert (p["lanes_per_row"], p["reg_tiles"], p["vec_elems"]) == (32, 2, 2)
    assert (p["block"], p["grid"]) == (256, 3 * SMS)            # 3 CTAs x 8 warps per SM
    assert (p["rows_per_slice"], p["rows_per_warp"], p["stages"]) == (16, 256, 3)
    assert p["reg_flavour"] == 5


def test_north_star_x64_f64_uses_grouped_lanes():
    p = plan(F64, L3D["rows"], L3D["nnz"], L3D["max_row"], 256, 64)
    assert (p["lanes_per_row"], p["reg_tiles"]) == (8, 4)        # four rows per warp side by side
    assert (p["block"], p["grid"]) == (768, SMS)                 # one CTA of 24 warps per SM
    assert (p["rows_per_slice"], p["rows_per_warp"], p["stages"]) == (16, 256, 2)
    # uneven rows (hub rows) stay on the warp-per-row stream
    q = plan(F64, L3D["rows"], L3D["nnz"], 60, 0, 64)
    assert (q["lanes_per_row"], q["reg_tiles"]) == (32, 1)


@pytest.mark.parametrize("dtype,n,lanes,tiles", [(F64, 32, 8, 2), (F32, 64, 8, 2), (F64, 16, 4, 2), (F32, 128, 8, 4)])
def test_narrow_rows_are_grouped_on_short_regular_rows(dtype, n, lanes, tiles):
    p = plan(dtype, L3D["rows"], L3D["nnz"], L3D["max_row"], 256, n)
    assert (p["lanes_per_row"], p["reg_tiles"]) == (lanes, tiles)
    assert p["rows_per_warp"] == 256


@pytest.mark.parametrize("dtype,n", [(F64, 8), (F32, 16)])
def test_64_byte_rows_walk_flat_streams_on_short_regular_rows(dtype, n):
    # four 128-bit lanes per row, 8 rows side by side, each lane group one flat entry stream over 8 rows of a 64-row slice
    p = plan(dtype, L3D["rows"], L3D["nnz"], L3D["max_row"], 256, n)
    assert (p["lanes_per_row"], p["reg_tiles"], p["reg_flavour"]) == (4, 1, 9)
    assert (p["rows_per_slice"], p["rows_per_warp"], p["stages"], p["block"]) == (64, 128, 2 if dtype == F64 else 3, 512)   # half a line per warp
    # long or uneven rows stay row by row; so do two-lane shapes
    assert plan(dtype, 1 << 20, 68_156_384, 65, 0, n)["reg_flavour"] == 1
    assert plan(dtype, L3D["rows"], L3D["nnz"], 60, 0, n)["reg_flavour"] == 1
    assert plan(dtype, L3D["rows"], L3D["nnz"], L3D["max_row"], 256, n // 2)["reg_flavour"] == 1
    # explicit request
    assert plan(dtype, 1 << 20, 68_156_384, 65, 0, n, reg_flavour=9)["reg_flavour"] == 9


def test_long_rows_stay_row_by_row_on_narrow_shapes():
    # band, half-bandwidth 32, x 32 f32 (when the row-block kernel is not used): 8 lanes, one tile, row by row
    p = plan(F32, 1 << 20, 68_156_384, 65, 0, 32)
    assert (p["lanes_per_row"], p["reg_tiles"]) == (8, 1)


def test_spmv_is_a_row_per_lane():
    p = plan(F64, 2048 ** 2, 20_963_328, 5, 2048, 1)
    assert (p["lanes_per_row"], p["reg_tiles"], p["vec_elems"]) == (1, 1, 1)


@pytest.mark.parametrize("line,want_slice", [(256, 16), (252, 12), (4096, 16), (250, 16), (100, 16)])
def test_rows_per_warp_follow_the_line_length_whatever_it_is(line, want_slice):
    """252: the slice is cut to a divisor of the line; 250 (no multiple of 4 divides it): short last slice, P = 250.
    A slice that does not divide the line used to push P to the next multiple (4096 -> 4104) and cost 60 %."""
    rows = line ** 3 if line <= 256 else line ** 2
    mean = 7 if line <= 256 else 5
    p = plan(F64, rows, rows * mean, mean, line, 128)
    assert p["rows_per_warp"] in (line, line // 2, line // 4), p
    assert line % p["rows_per_warp"] == 0
    assert p["rows_per_slice"] == want_slice, p


def test_small_row_blocks_trade_line_length_for_balance():
    # a 1/8 row block of the headline matrix: half a line per warp loses less to wave quantisation
    p = plan(F64, L3D["rows"] // 8, L3D["nnz"] // 8, 7, 256, 128)
    assert p["rows_per_warp"] == 128


def test_tuning_overrides_and_slices_that_do_not_fit():
    p = plan(F64, L3D["rows"], L3D["nnz"], 7, 256, 128, rows_per_slice=32, stages=2, rows_per_warp=96, warps_per_cta=4)
    assert (p["rows_per_slice"], p["stages"], p["rows_per_warp"], p["block"]) == (32, 2, 96, 128)
    # rows too long for any stage: the unstaged variant (capacity 0), never the grouped shapes
    q = plan(F64, 100_000, 100_000 * 50, 40_000, 0, 64)
    assert q["capacity"] == 0 and q["reg_flavour"] == 0 and (q["lanes_per_row"], q["reg_tiles"]) == (32, 1)


def test_few_rows_shrink_the_cta():
    p = plan(F64, 20_000, 140_000, 7, 0, 64)
    assert p["block"] < 768 and p["grid"] >= 100


def line_length(cols, diag):
    arr = (C.c_uint32 * len(cols))(*cols)
    return _lib.lib().bsm_line_length_of_row(arr, len(cols), diag)


def test_line_length_from_a_row():
    nx, i = 4096, 4096 * 1000
    assert line_length([i - nx, i - 1, i, i + 1, i + nx], i) == nx            # interior row of a 5-point stencil
    assert line_length([i - nx, i, i + 1, i + nx], i) == nx                   # x = 0 boundary row (the median column is i + 1)
    assert line_length([i - nx, i - 1, i, i + nx], i) == nx                   # x = nx - 1 boundary row
    assert line_length([i, i + 1, i + nx], i) == nx                           # first line
    p = 256
    j = 65536 * 100 + 256 * 7
    assert line_length([j - p * p, j - p, j, j + 1, j + p, j + p * p], j) == p  # 7-point, x = 0
    box = [j + dz * p * p + dy * p + dx for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    assert line_length(sorted(box), j) == p                                   # 27-point: distances p-1, p, p+1 -> p
    assert line_length([j - 1, j, j + 1], j) == 0                             # tridiagonal: no line
    # row blocks of a partitioned matrix keep global columns: the diagonal is local row + row offset
    assert line_length([i - nx, i - 1, i, i + 1, i + nx], (i - 12345) + 12345) == nx


# ---- invariants of every plan (what launch_spmm_rows and the kernel rely on), over random inputs ---------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=400, deadline=None)
@given(dtype=st.sampled_from([F32, F64]),
       rows=st.integers(1, 1 << 26),
       mean=st.floats(0.1, 300.0),
       spread=st.floats(1.0, 60.0),
       stride=st.sampled_from([0, 0, 16, 37, 50, 100, 250, 252, 256, 1000, 4096, 16384]),
       n=st.integers(1, 600),
       tune=st.fixed_dictionaries({}, optional={
           "rows_per_slice": st.sampled_from([4, 8, 12, 16, 32, 64, 256]), "stages": st.integers(1, 8),
           "warps_per_cta": st.sampled_from([1, 2, 3, 4, 8, 16, 24]), "rows_per_warp": st.sampled_from([4, 16, 33, 50, 128, 251, 512]),
           "reg_flavour": st.integers(1, 8), "lanes_per_row": st.sampled_from([4, 8, 16, 32]), "prefer_wide_rows": st.sampled_from([0, 1]),
           "col_tile": st.sampled_from([8, 16, 32, 64, 128])}))
def test_every_plan_is_launchable(dtype, rows, mean, spread, stride, n, tune):
    nnz = max(1, min(int(rows * mean), (1 << 32) - 64))
    max_row = max(1, min(int(max(mean, 1.0) * spread) + 1, nnz))
    t = _lib.Tuning()
    for k, v in tune.items():
        setattr(t, k, v)
    info = _lib.LaunchInfo()
    status = _lib.lib().bsm_plan_vector(dtype, rows, nnz, max_row, stride, n, C.byref(t), SMS, SMEM, C.byref(info))
    if status != 0:                                    # only one legitimate refusal: a user-fixed slice that cannot fit
        assert b"does not fit shared memory" in _lib.lib().bsm_last_error_string() and "rows_per_slice" in tune
        return
    p = info.as_dict()
    s = 4 if dtype == F32 else 8
    G, NT, V, R, P = p["lanes_per_row"], p["reg_tiles"], p["vec_elems"], p["rows_per_slice"], p["rows_per_warp"]
    assert G in (1, 2, 4, 8, 16, 32) and NT in (1, 2, 4) and V * s <= 16 and V >= 1
    width = min(n, p["col_tile"])
    assert G * NT * V >= width                          # one pass covers its columns
    assert R >= 4 and R % 4 == 0 and R % max(1, 32 // G) == 0
    assert P >= R
    if not (G == 32 or NT > 1):
        assert P % R == 0                               # the row-by-row shapes need whole slices
    assert 32 <= p["block"] <= 768 and p["block"] % 32 == 0
    assert 1 <= p["stages"] <= 8
    assert p["smem_bytes"] <= SMEM - 1024
    if p["capacity"]:                                   # staged: any R consecutive rows fit (+ aligned start)
        assert p["capacity"] % 4 == 0 and p["capacity"] >= R * max_row + 3
    else:
        assert p["reg_flavour"] == 0                    # unstaged variant
    warps = p["block"] // 32
    supers = -(-rows // (warps * P))
    assert 1 <= p["grid"] <= min(supers, 4 * SMS)
    assert p["col_tile"] <= n and p["passes"] >= -(-n // p["col_tile"])   # col_tile = width of the first pass


@pytest.mark.parametrize("dtype,n,first,passes", [(F32, 129, 128, 2), (F64, 129, 128, 2), (F64, 255, 254, 2), (F32, 258, 256, 2),
                                                   (F32, 513, 512, 2), (F64, 300, 256, 2), (F32, 300, 300, 1), (F32, 130, 128, 2), (F64, 130, 130, 1)])
def test_pass_width_when_alignment_forces_narrow_vectors(dtype, n, first, passes):
    """129 f32 columns are 129 one-element lanes — more than four register tiles of 32 lanes hold. The pass is cut to a
    width the 16-byte vectors divide (128) and the rest goes to the next pass; before this the last column was dropped.
    (f32 lanes are 4 elements or 1: 130 f32 columns are one-element lanes too — there are no 2-element f32 kernels.)"""
    p = plan(dtype, 1_000_000, 7_000_000, 7, 0, n)
    assert (p["col_tile"], p["passes"]) == (first, passes)
    assert p["lanes_per_row"] * p["reg_tiles"] * p["vec_elems"] >= first
