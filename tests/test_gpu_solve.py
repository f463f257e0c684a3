"""GPU: forward / backward substitution (bsm_forward_substitution / bsm_backward_substitution) against the reference's own
f32 KATs (lib.rs:73-137), against a statement-by-statement restatement of lib.rs:28-65 on adversarial factors, and against
the banded oracle at config-5 scale — all BITWISE, like the reference's assert_eq! on f32."""
import numpy as np
import pytest

from basic_sparse_matrix_b200 import Csr, Dense, MatErr, MatError, _lib
from helpers import assert_bitwise

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def gpu():
    from basic_sparse_matrix_b200 import gpu as g
    if g.device_count() < 1:
        pytest.skip("no CUDA device")
    g.init(0)
    return g


def dense_rows_to_csr(rows, dtype=f32):
    """Csr::from_data(rows): zeros are not stored (sparse.rs:193-203 -> insert)."""
    return Csr.from_data(rows, dtype)


def test_forward_substitution_kat(gpu):          # lib.rs:73-93
    from basic_sparse_matrix_b200 import solve
    l = dense_rows_to_csr([[5, 0, 0], [8, 2, 0], [3, 7, 1]])
    y = solve.forward_substitution(l, Dense.from_data([[7, 3, 1]], f32))
    assert_bitwise(np.asarray(y.get_col(0)), np.array([f32(7.0) / f32(5.0), -4.1, 25.5], f32), "forward_substitution_test_0")


def test_backward_substitution_kat(gpu):         # lib.rs:95-115
    from basic_sparse_matrix_b200 import solve
    l_star = dense_rows_to_csr([[7, 1, 8], [0, 2, 3], [0, 0, 5]])
    x = solve.backward_substitution(l_star, Dense.from_data([[1, 7, 3]], f32))
    assert_bitwise(np.asarray(x.get_col(0)), np.array([f32(-32.0) / f32(35.0), 2.6, 0.6], f32), "backward_substitution_test_0")


def test_solve_kat_with_the_oracle_factor(gpu):  # lib.rs:117-137: x_ref = [0.625, -0.1, 2.6999998, 0.5]
    from basic_sparse_matrix_b200 import solve
    from oracle import ref_solve
    a = np.array([[8, 0, 0, 0], [0, 7, 1, 0], [0, 1, 3, 0], [0, 0, 0, 2]], f32)
    l_band = ref_solve.cholesky_band(ref_solve.dense_to_band(a, 3))        # factorisation: CPU (out of scope on the GPU)
    l = Csr.from_raw_parts((4, 4), *ref_solve.band_to_csr_lower(l_band))
    l_star = Csr.from_raw_parts((4, 4), *ref_solve.band_to_csr_upper(l_band))
    x = solve.solve_with_factor(l, l_star, Dense.from_data([[5, 2, 8, 1]], f32))
    assert_bitwise(np.asarray(x.get_col(0)), np.array([0.625, -0.1, 2.6999998, 0.5], f32), "solve_test")


def test_dimension_and_empty_row_errors(gpu):
    from basic_sparse_matrix_b200 import solve
    l = dense_rows_to_csr([[5, 0, 0], [8, 2, 0], [3, 7, 1]])
    with pytest.raises(MatError) as e:
        solve.forward_substitution(l, Dense.from_data([[7, 3]], f32))
    assert e.value.kind == MatErr.IncorrectDimensions
    empty = dense_rows_to_csr([[5, 0, 0], [0, 0, 0], [3, 7, 1]])          # row 1 has no stored entry: the reference panics
    with pytest.raises(_lib.BsmError):
        solve.forward_substitution(empty, Dense.from_data([[7, 3, 1]], f32))


def _random_factor(rng, n, dtype, far=False, upper=False):
    """Adversarial triangular factor as raw Csr parts: ragged rows, unsorted off-diagonal columns, duplicates, entries on
    the wrong side of the diagonal (they read the still-default 0.0), and — with far=True — dependencies further back than
    the kernel's shared-memory ring. The diagonal sits where the reference expects it (last / first stored entry)."""
    vals, cols, ri = [], [], [0]
    for r in range(n):
        k = int(rng.integers(0, 7))
        if upper:
            cand = rng.integers(r, min(n, r + (900 if far else 40)), size=k) if r + 1 < n else np.array([], int)
        else:
            cand = rng.integers(max(0, r - (900 if far else 40)), r + 1, size=k) if r > 0 else np.array([], int)
        off = [int(c) for c in cand]
        if rng.random() < 0.1 and n > 4:                       # an entry on the wrong side
            off.append(int(rng.integers(0, n)))
        offv = [float(rng.uniform(-0.2, 0.2)) for _ in off]
        diag = float(rng.uniform(1.0, 2.0))
        if upper:
            cols += [r] + off
            vals += [diag] + offv
        else:
            cols += off + [r]
            vals += offv + [diag]
        ri.append(len(cols))
    return np.array(vals, dtype), np.array(cols, np.uint64), np.array(ri, np.uint64)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("far", [False, True])
def test_substitutions_bitwise_on_adversarial_factors(gpu, dtype, far):
    from oracle import ref_solve
    rng = np.random.default_rng(77)
    n = 1500 if far else 400
    for nrhs in (1, 5, 33, 70):
        b_cols = rng.uniform(0.5, 1.5, (nrhs, n)).astype(dtype)
        b_dev = gpu.DeviceDense.from_rowmajor(np.ascontiguousarray(b_cols.T))
        v, ci, ri = _random_factor(rng, n, dtype, far=far)
        with gpu.DeviceCsr.from_host(Csr.from_raw_parts((n, n), v, ci, ri)) as l, l.forward_substitution(b_dev) as y:
            assert_bitwise(y.to_rowmajor().T, ref_solve.forward_csr(v, ci, ri, b_cols), f"forward n={n} nrhs={nrhs} far={far}")
        v, ci, ri = _random_factor(rng, n, dtype, far=far, upper=True)
        with gpu.DeviceCsr.from_host(Csr.from_raw_parts((n, n), v, ci, ri)) as ls, ls.backward_substitution(b_dev) as x:
            assert_bitwise(x.to_rowmajor().T, ref_solve.backward_csr(v, ci, ri, b_cols), f"backward n={n} nrhs={nrhs} far={far}")
        b_dev.close()


def test_rows_longer_than_a_stage_are_read_unstaged(gpu):
    """A dense lower-triangular factor: rows of up to 1200 entries do not fit the TMA stage -> col_idx / values from global."""
    from oracle import ref_solve
    rng = np.random.default_rng(5)
    n = 1200
    dense = np.tril(rng.uniform(-0.01, 0.01, (n, n))).astype(f32)
    dense[np.arange(n), np.arange(n)] = rng.uniform(1.0, 2.0, n).astype(f32)
    rows_i, cols_i = np.nonzero(dense)                       # row-major order = insertion order of Csr::from_data
    ri = np.zeros(n + 1, np.uint64)
    np.cumsum(np.bincount(rows_i, minlength=n), out=ri[1:])
    v, ci = dense[rows_i, cols_i], cols_i.astype(np.uint64)
    m = Csr.from_raw_parts((n, n), v, ci, ri)
    b_cols = rng.uniform(0.5, 1.5, (3, n)).astype(f32)
    with gpu.DeviceCsr.from_host(m) as l, gpu.DeviceDense.from_rowmajor(np.ascontiguousarray(b_cols.T)) as b, l.forward_substitution(b) as y:
        assert_bitwise(y.to_rowmajor().T, ref_solve.forward_csr(v, ci, ri, b_cols), "dense factor, unstaged")


@pytest.mark.parametrize("n_rows", [1 << 14])
def test_band_solve_matches_the_reference_solve_bitwise(gpu, n_rows):
    """BASELINE config 5 at 2^14 rows: factor on the CPU (banded restatement of cholesky_decomp, pinned on the f32 KATs),
    BOTH substitutions on the GPU; X must equal the reference's `solve` bit for bit (32 right-hand sides)."""
    from basic_sparse_matrix_b200 import gen
    from oracle import ref_solve
    hb, nrhs = 32, 32
    a_band = ref_solve.spd_band(n_rows, hb)
    b_cols = gen.dense_rows(n_rows, nrhs, 6, gen.MODE_REAL, 0.5, f32).T.copy()
    x_ref = ref_solve.solve_band(a_band, b_cols)
    l_band = ref_solve.cholesky_band(a_band)
    with gpu.DeviceCsr.from_host(Csr.from_raw_parts((n_rows, n_rows), *ref_solve.band_to_csr_lower(l_band))) as l, \
            gpu.DeviceCsr.from_host(Csr.from_raw_parts((n_rows, n_rows), *ref_solve.band_to_csr_upper(l_band))) as ls, \
            gpu.DeviceDense.generate(n_rows, nrhs, seed=6, mode=gen.MODE_REAL, offset=0.5, dtype=f32) as b:
        with l.forward_substitution(b) as y, ls.backward_substitution(y) as x:
            assert_bitwise(x.to_rowmajor().T, x_ref, "device substitutions vs reference solve")


# ---- proper band factors: the specialised kernels (one solver warp, staging warps) ------------------------------------------
def _band_factor(rng, n, hb, dtype, upper=False, ragged_row=None):
    """Raw Csr parts of a proper lower (diagonal last) / upper (diagonal first) band factor of half-bandwidth hb."""
    vals, cols, ri = [], [], [0]
    for r in range(n):
        off_cols = list(range(r + 1, min(n, r + hb + 1))) if upper else list(range(max(0, r - hb), r))
        if ragged_row is not None and r == ragged_row and off_cols:
            off_cols = off_cols[1:] if not upper else off_cols[:-1]      # one entry short: no longer a proper band
        offv = rng.uniform(-0.05, 0.05, len(off_cols)).tolist()
        diag = float(rng.uniform(1.0, 2.0))
        cols += ([r] + off_cols) if upper else (off_cols + [r])
        vals += ([diag] + offv) if upper else (offv + [diag])
        ri.append(len(cols))
    return np.array(vals, dtype), np.array(cols, np.uint64), np.array(ri, np.uint64)


def _substitute_np(v, ci, ri, b_cols, upper):
    """lib.rs:35-42 / 56-60 with the right-hand sides vectorised: one numpy operation per stored entry, all columns at once (every
    operation separately rounded in the array dtype, in stored order — the same arithmetic as ref_solve.forward_csr / backward_csr,
    which the small cases below are also checked against)."""
    n = b_cols.shape[1]
    x = np.zeros_like(b_cols)
    ci = ci.astype(np.int64)
    ri = ri.astype(np.int64)
    for r in (range(n - 1, -1, -1) if upper else range(n)):
        lx = np.zeros(b_cols.shape[0], b_cols.dtype)
        s, e = ri[r], ri[r + 1]
        for k in (range(s + 1, e) if upper else range(s, e)):
            if upper or ci[k] != r:
                lx = lx + v[k] * x[:, ci[k]]
        x[:, r] = (b_cols[:, r] - lx) / (v[s] if upper else v[e - 1])
    return x


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("hb", [8, 16, 32])
def test_band_kernels_bitwise(gpu, dtype, hb, monkeypatch):
    """Proper band factors take the specialised kernels: checked against the restatement of lib.rs:28-65 AND against the general
    kernel (BSM_SOLVE_GENERAL=1), for row counts around the batch / group sizes and ragged numbers of right-hand sides."""
    from oracle import ref_solve
    rng = np.random.default_rng(1000 + hb)
    for n, nrhs in ((4 * hb, 1), (4 * hb + 3, 33), (4 * hb + 16, 5), (1000 + hb, 32), (2049, 45)):
        b_cols = rng.uniform(0.5, 1.5, (nrhs, n)).astype(dtype)
        with gpu.DeviceDense.from_rowmajor(np.ascontiguousarray(b_cols.T)) as b_dev:
            for upper in (False, True):
                v, ci, ri = _band_factor(rng, n, hb, dtype, upper=upper)
                want = _substitute_np(v, ci, ri, b_cols, upper)
                if n <= 4 * hb + 3:
                    small = (ref_solve.backward_csr if upper else ref_solve.forward_csr)(v, ci, ri, b_cols)
                    assert_bitwise(want, small, "vectorised restatement vs statement-by-statement restatement")
                with gpu.DeviceCsr.from_host(Csr.from_raw_parts((n, n), v, ci, ri)) as l:
                    bs = l.band_structure()
                    assert bs == ({"lower_hb": -1, "upper_hb": hb} if upper else {"lower_hb": hb, "upper_hb": -1}), bs
                    sub = l.backward_substitution if upper else l.forward_substitution
                    with sub(b_dev) as x:
                        got = x.to_rowmajor().T
                    monkeypatch.setenv("BSM_SOLVE_GENERAL", "1")
                    with sub(b_dev) as x:
                        general = x.to_rowmajor().T
                    monkeypatch.delenv("BSM_SOLVE_GENERAL")
                assert_bitwise(got, want, f"band kernel hb={hb} n={n} nrhs={nrhs} upper={upper}")
                assert_bitwise(general, want, f"general kernel hb={hb} n={n} nrhs={nrhs} upper={upper}")


def test_band_probe_rejects_everything_else(gpu):
    rng = np.random.default_rng(3)
    n, hb = 300, 16
    cases = {
        "ragged row": _band_factor(rng, n, hb, f32, ragged_row=150),
        "half-bandwidth the kernels are not built for": _band_factor(rng, n, 12, f32),
        "adversarial": _random_factor(rng, n, f32),
    }
    b_cols = rng.uniform(0.5, 1.5, (3, n)).astype(f32)
    from oracle import ref_solve
    with gpu.DeviceDense.from_rowmajor(np.ascontiguousarray(b_cols.T)) as b_dev:
        for name, (v, ci, ri) in cases.items():
            with gpu.DeviceCsr.from_host(Csr.from_raw_parts((n, n), v, ci, ri)) as l:
                bs = l.band_structure()
                assert bs["upper_hb"] == -1 and bs["lower_hb"] == (12 if "built for" in name else -1), (name, bs)
                with l.forward_substitution(b_dev) as y:
                    assert_bitwise(y.to_rowmajor().T, ref_solve.forward_csr(v, ci, ri, b_cols), name)
    # special values travel exactly like in the reference: an infinite right-hand side entry turns the rows below into NaN / inf
    v, ci, ri = _band_factor(rng, 200, 8, f32)
    b_cols = rng.uniform(0.5, 1.5, (2, 200)).astype(f32)
    b_cols[0, 37] = np.inf
    b_cols[1, 120] = np.nan
    with gpu.DeviceDense.from_rowmajor(np.ascontiguousarray(b_cols.T)) as b_dev, \
            gpu.DeviceCsr.from_host(Csr.from_raw_parts((200, 200), v, ci, ri)) as l, l.forward_substitution(b_dev) as y:
        with np.errstate(all="ignore"):
            want = _substitute_np(v, ci, ri, b_cols, False)
        got = y.to_rowmajor().T
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(want)], want[~np.isnan(want)])


@pytest.mark.parametrize("upper", [False, True])
def test_band_kernels_zero_and_special_right_hand_sides(gpu, upper):
    """Zero numerators (a zero column, -0.0 entries) stay on the band kernels' division shortcut and keep the sign of their zeros;
    infinities, NaN, subnormals and solutions that decay towards them make the library fall back to the general kernel —
    bit-identical either way."""
    rng = np.random.default_rng(17)
    n, hb = 700, 32
    v, ci, ri = _band_factor(rng, n, hb, f32, upper=upper)
    b_cols = rng.uniform(0.5, 1.5, (6, n)).astype(f32)
    b_cols[0, :] = 0.0                         # a zero column: every numerator is +-0
    b_cols[1, ::2] = -0.0
    b_cols[2, 5] = -0.0
    with np.errstate(all="ignore"):
        want = _substitute_np(v, ci, ri, b_cols, upper)
    with gpu.DeviceDense.from_rowmajor(np.ascontiguousarray(b_cols.T)) as b_dev, \
            gpu.DeviceCsr.from_host(Csr.from_raw_parts((n, n), v, ci, ri)) as l:
        sub = l.backward_substitution if upper else l.forward_substitution
        with gpu.DeviceDense.from_rowmajor(np.ascontiguousarray(rng.uniform(0.5, 1.5, (n, 6)).astype(f32))) as b_plain:
            sub(b_plain).close()                                  # (the probes of a new handle run here)
            c0 = gpu.kernel_launch_count()
            sub(b_plain).close()
            plain_launches = gpu.kernel_launch_count() - c0       # what one substitution on the band kernel launches
        c0 = gpu.kernel_launch_count()
        with sub(b_dev) as x:
            launches = gpu.kernel_launch_count() - c0
            assert_bitwise(x.to_rowmajor().T, want, "zero numerators")
        assert launches == plain_launches, "zero numerators must not leave the band kernel"
    # (a unit vector: the solution decays below 2^-90 a few dozen rows after the one, then through the subnormals to zero)
    for special in (float(np.inf), float(np.nan), f32(1e-42), f32(3e38), "unit vector"):
        b2 = b_cols.copy()
        if special == "unit vector":
            b2[3, :] = 0.0
            b2[3, 350] = 1.0
        else:
            b2[3, 350] = special
        with np.errstate(all="ignore"):
            want = _substitute_np(v, ci, ri, b2, upper)
        with gpu.DeviceDense.from_rowmajor(np.ascontiguousarray(b2.T)) as b_dev, \
                gpu.DeviceCsr.from_host(Csr.from_raw_parts((n, n), v, ci, ri)) as l:
            sub = l.backward_substitution if upper else l.forward_substitution
            sub(b_dev).close()
            c0 = gpu.kernel_launch_count()
            with sub(b_dev) as x:
                # (the subnormal right-hand side entry disappears in b - l_x: that numerator is an ordinary number)
                extra = gpu.kernel_launch_count() - c0 - plain_launches
                if isinstance(special, float):                     # inf, nan
                    assert extra == 1, f"{special}: expected the general kernel after the band kernel"
                assert extra in (0, 1)
                got = x.to_rowmajor().T
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan) and np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32)), special
