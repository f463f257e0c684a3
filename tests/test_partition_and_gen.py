"""CPU: host-side sharding logic (bsm_partition_rows is a pure host function of the C ABI) and
the numpy generators that mirror csrc/gen.cu."""
import numpy as np
import pytest
import scipy.sparse as sp

from basic_sparse_matrix_b200 import gen
from basic_sparse_matrix_b200.gpu import partition_rows


def test_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(1)
    lens = rng.poisson(5, size=10_000)
    lens[1234] = 20_000                       # one hub row
    ri = np.zeros(len(lens) + 1, np.uint64)
    np.cumsum(lens, out=ri[1:])
    for parts in (1, 2, 3, 4, 8):
        b = partition_rows(ri, parts)
        assert b[0] == 0 and b[-1] == len(lens) and np.all(np.diff(b.astype(np.int64)) >= 0)
        nnz = np.diff(ri[b.astype(np.int64)].astype(np.int64))
        assert nnz.sum() == ri[-1]
        # every part within one max-row of the ideal share
        assert np.all(np.abs(nnz - ri[-1] / parts) <= lens.max() + 1)


def test_partition_degenerate():
    ri = np.zeros(6, np.uint64)               # 5 empty rows
    assert partition_rows(ri, 4).tolist() == [0, 1, 2, 3, 5]
    ri = np.array([0, 0, 0, 100], np.uint64)  # everything in the last row: cannot be split
    b = partition_rows(ri, 4)
    assert b[0] == 0 and b[-1] == 3


def test_laplacian_matches_scipy():
    nx, ny, nz = 7, 5, 4
    v, ci, ri, dims = gen.laplacian(nx, ny, nz)
    a = sp.csr_matrix((v, ci.astype(np.int64), ri.astype(np.int64)), shape=dims)
    ex = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(nx, nx))
    ey = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(ny, ny))
    ez = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(nz, nz))
    ref = (sp.kron(sp.eye(nz), sp.kron(sp.eye(ny), ex)) + sp.kron(sp.eye(nz), sp.kron(ey, sp.eye(nx)))
           + sp.kron(ez, sp.kron(sp.eye(ny), sp.eye(nx))))
    assert abs(a - ref).sum() == 0
    # columns ascending inside every row; row slices are consistent with the full matrix
    for r in range(dims[0]):
        c = ci[int(ri[r]):int(ri[r + 1])]
        assert np.all(np.diff(c.astype(np.int64)) > 0)
    v2, ci2, ri2, d2 = gen.laplacian(nx, ny, nz, 30, 90)
    assert np.array_equal(v2, v[int(ri[30]):int(ri[90])]) and np.array_equal(ci2, ci[int(ri[30]):int(ri[90])])
    assert np.array_equal(ri2, ri[30:91] - ri[30])
    assert gen.laplacian_row_counts(nx, ny, nz).tolist() == np.diff(ri.astype(np.int64)).tolist()


def test_config_sizes_closed_form():
    # 2048^2 5-point -> 20 963 328 nnz; 256^3 7-point -> 117 047 296 nnz
    assert 5 * 2048 * 2048 - 4 * 2048 == 20_963_328
    assert 7 * 256 ** 3 - 6 * 256 ** 2 == 117_047_296
    assert int(gen.laplacian_row_counts(64, 64, 1).sum()) == 5 * 64 * 64 - 4 * 64


def test_band_is_spd_shaped():
    v, ci, ri, dims = gen.band(50, 4)
    a = sp.csr_matrix((v, ci.astype(np.int64), ri.astype(np.int64)), shape=dims).toarray()
    assert np.allclose(a, a.T)
    off = np.abs(a).sum(axis=1) - np.abs(np.diag(a))
    assert np.all(np.diag(a) > off)           # strictly diagonally dominant
    assert np.linalg.eigvalsh(a).min() > 0
    assert a[10, 12] == -1.0 / 3.0 and a[10, 15] == 0.0


def test_rmat_shape():
    v, ci, ri, dims = gen.rmat(8, 4000, seed=3)
    assert dims == (256, 256) and ri[-1] == 4000 and len(v) == 4000
    assert np.all(np.diff(ri.astype(np.int64)) >= 0) and ci.max() < 256
    lens = np.diff(ri.astype(np.int64))
    assert lens.max() > 8 * lens.mean()       # power-law: a hub row far above the mean
    assert np.all((v * 256) == np.round(v * 256)) and v.min() > 0


def test_hash_is_stable():
    # pinned values: the device generator (gen.cu) must reproduce these
    h = gen.hash_u64(5, np.arange(3))
    assert h.dtype == np.uint64
    assert gen.hash_u64(5, 0) == h[0] and len(set(h.tolist())) == 3
    d = gen.dense_rows(4, 3, seed=5, mode=gen.MODE_EXACT)
    assert d.shape == (4, 3) and np.all((d * 1024) == np.round(d * 1024)) and d.min() >= 0 and d.max() < 1
    assert np.array_equal(gen.dense_rows(4, 3, 5, gen.MODE_EXACT, row_ids=[2]), d[2:3])


def test_bench_as_written_piles_into_last_rows():
    a, x = gen.bench_as_written(2000)
    lens = np.diff(a.row_index.astype(np.int64))
    assert a.get_dims().rows == 1000 and x.get_dims().cols == 10 and x.get_dims().rows == 1000
    assert lens[-1] > 0.8 * lens.sum()        # nearly everything lands in the last row (sparse.rs:237-250)
    assert (lens > 0).sum() < 20
