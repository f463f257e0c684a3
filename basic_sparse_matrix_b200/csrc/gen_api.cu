// gen_api.cu — C ABI of the synthetic workloads generated directly in HBM (kernels in gen.cu).
#include <string>

#include "bsm_internal.h"

namespace bsm {

template <typename CountFn, typename FillFn>
static int gen_counted(int dtype, uint64_t rows, uint64_t cols, uint64_t max_per_row, uint64_t row_begin, CountFn count_fn, FillFn fill_fn, bsm_csr **out)
{
    BSM_TRY(ensure_init());
    if (!out) return fail(BSM_ERR_INVALID_ARGUMENT, "gen: null out");
    if (dtype != BSM_F32 && dtype != BSM_F64) return fail(BSM_ERR_DTYPE_MISMATCH, "gen: dtype must be f32 or f64");
    if (rows >= 0xFFFFFFF0ull) return fail(BSM_ERR_INDEX_OVERFLOW, "gen: too many rows");
    // the row counts are scanned in u32: refuse anything whose entry count could wrap
    if ((__uint128_t)rows * max_per_row >= 0xFFFFFFF0ull) return fail(BSM_ERR_INDEX_OVERFLOW, "gen: nnz would not fit the device's u32 indices");
    cudaStream_t sm = rt().stream;
    uint32_t *counts = nullptr;
    bsm_csr *a = nullptr;
    int st = [&]() -> int {
        BSM_TRY(tmp_alloc((void **)&counts, (rows + 1) * 4 + 16));
        BSM_CUDA(cudaMemsetAsync(counts, 0, (rows + 1) * 4 + 16, sm));
        BSM_TRY(count_fn(counts, sm));
        BSM_TRY(exclusive_scan_u32(counts, counts, rows + 1, sm));
        uint32_t nnz = 0;
        BSM_CUDA(cudaMemcpyAsync(&nnz, counts + rows, 4, cudaMemcpyDeviceToHost, sm));
        BSM_CUDA(cudaStreamSynchronize(sm));
        BSM_TRY(alloc_csr(dtype, rows, cols, nnz, &a));
        a->row_offset = row_begin;
        BSM_CUDA(cudaMemcpyAsync(a->row_ptr, counts, (rows + 1) * 4, cudaMemcpyDeviceToDevice, sm));
        BSM_TRY(fill_fn(a, sm));
        return compute_stats(a, false);
    }();
    tmp_free(counts);
    if (st != BSM_OK) {
        if (a) bsm_csr_free(a);
        return st;
    }
    *out = a;
    return BSM_OK;
}

}  // namespace bsm

using namespace bsm;

extern "C" {

// ---- generators ------------------------------------------------------------------------------------
int bsm_gen_dense(int dtype, uint64_t rows, uint64_t cols, uint64_t seed, int mode, double offset, bsm_dense **out)
{
    bsm_dense *d = nullptr;
    BSM_TRY(dense_alloc(dtype, rows, cols, &d));
    int st = launch_gen_dense(dtype, d->data, rows, cols, d->ld, seed, mode, offset, rt().stream);
    if (st != BSM_OK) {
        bsm_dense_free(d);
        return st;
    }
    *out = d;
    return BSM_OK;
}

int bsm_gen_laplacian(int dtype, uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_begin, uint64_t row_end, bsm_csr **out)
{
    if (nx == 0 || ny == 0 || nz == 0) return fail(BSM_ERR_INVALID_ARGUMENT, "gen_laplacian: empty grid");
    const uint64_t n = nx * ny * nz;
    if (row_begin > row_end || row_end > n) return fail(BSM_ERR_INVALID_ARGUMENT, "gen_laplacian: bad row range");
    return gen_counted(
        dtype, row_end - row_begin, n, 1 + 2 * ((nx > 1) + (ny > 1) + (nz > 1)), row_begin,
        [&](uint32_t *counts, cudaStream_t sm) { return launch_laplacian_counts(nx, ny, nz, row_begin, row_end, counts, sm); },
        [&](bsm_csr *a, cudaStream_t sm) {
            return launch_laplacian_fill(dtype, nx, ny, nz, row_begin, row_end, a->row_ptr, a->col_idx, a->vals, sm);
        },
        out);
}

int bsm_gen_band(int dtype, uint64_t n, uint64_t hb, uint64_t row_begin, uint64_t row_end, bsm_csr **out)
{
    if (n == 0 || row_begin > row_end || row_end > n) return fail(BSM_ERR_INVALID_ARGUMENT, "gen_band: bad arguments");
    return gen_counted(
        dtype, row_end - row_begin, n, 2 * hb + 1, row_begin,
        [&](uint32_t *counts, cudaStream_t sm) { return launch_band_counts(n, hb, row_begin, row_end, counts, sm); },
        [&](bsm_csr *a, cudaStream_t sm) {
            return launch_band_fill(dtype, n, hb, row_begin, row_end, a->row_ptr, a->col_idx, a->vals, sm);
        },
        out);
}

int bsm_gen_rmat(int dtype, int scale, uint64_t edges, double pa, double pb, double pc, uint64_t seed, int mode, bsm_csr **out)
{
    BSM_TRY(ensure_init());
    if (!out) return fail(BSM_ERR_INVALID_ARGUMENT, "gen_rmat: null out");
    if (dtype != BSM_F32 && dtype != BSM_F64) return fail(BSM_ERR_DTYPE_MISMATCH, "gen_rmat: dtype must be f32 or f64");
    if (scale < 1 || scale > 31) return fail(BSM_ERR_INVALID_ARGUMENT, "gen_rmat: scale must be in [1,31]");
    const uint64_t rows = 1ull << scale;
    bsm_csr *a = nullptr;
    BSM_TRY(alloc_csr(dtype, rows, rows, edges, &a));
    int st = gen_rmat_device(dtype, scale, edges, pa, pb, pc, seed, mode, a->row_ptr, a->col_idx, a->vals, rt().stream);
    if (st == BSM_OK) st = compute_stats(a, false);
    if (st != BSM_OK) {
        bsm_csr_free(a);
        return st;
    }
    *out = a;
    return BSM_OK;
}

}  // extern "C"
