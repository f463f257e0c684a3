// handles.cu — the operand types of the path on the device: bsm_csr (Csr<T>, src/sparse.rs:68-78) and bsm_dense
// (Dense<T>, src/dense.rs:4-9): allocation, upload / download with format conversion (usize -> u32 narrowing,
// column-major <-> row-major), per-matrix statistics, borrowed and IPC views.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "bsm_internal.h"
#include "line_length.h"

namespace bsm {

// leading dimension of library-owned dense buffers: rows start 16-byte aligned whenever a row is
// at least 16 bytes, so lanes can use 128-bit loads
uint64_t default_ld(uint64_t cols, int dtype)
{
    const uint64_t v = 16 / dtype_size(dtype);
    if (cols <= 1) return cols ? cols : 1;
    return round_up(cols, v);
}

int alloc_csr(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, bsm_csr **out)
{
    if (rows >= 0xFFFFFFF0ull || cols >= 0xFFFFFFF0ull || nnz >= 0xFFFFFFF0ull)
        return fail(BSM_ERR_INDEX_OVERFLOW, "rows, cols and nnz must fit the device's u32 indices");
    bsm_csr *a = new bsm_csr();
    a->dtype = dtype;
    a->rows = rows;
    a->cols = cols;
    a->nnz = nnz;
    const size_t s = dtype_size(dtype);
    int st = dev_alloc(&a->vals, (pad4(nnz) + 4) * s, &a->pooled);
    if (st == BSM_OK) st = dev_alloc((void **)&a->col_idx, (pad4(nnz) + 4) * 4, &a->pooled);
    if (st == BSM_OK) st = dev_alloc((void **)&a->row_ptr, (pad4(rows + 1) + 4) * 4, &a->pooled);
    if (st != BSM_OK) {
        bsm_csr_free(a);
        return st;
    }
    // defined padding (TMA over-reads up to the next multiple of 4 entries)
    cudaStream_t sm = rt().stream;
    cudaMemsetAsync((char *)a->vals + nnz * s, 0, (pad4(nnz) + 4 - nnz) * s, sm);
    cudaMemsetAsync(a->col_idx + nnz, 0, (pad4(nnz) + 4 - nnz) * 4, sm);
    cudaMemsetAsync(a->row_ptr + rows + 1, 0, (pad4(rows + 1) + 4 - rows - 1) * 4, sm);
    *out = a;
    return BSM_OK;
}

// Statistics of a device Csr — longest row, column range, row_ptr sanity, dominant stencil line length — from ONE
// kernel (csr_stats_kernel, convert.cu) and one readback. `scratch` (kStatWords u32, prepared by stats_scratch_alloc)
// may already carry the flags of the upload's narrowing kernels, so an upload synchronises exactly once.
static int stats_scratch_alloc(uint32_t **d)
{
    BSM_TRY(tmp_alloc((void **)d, kStatWords * 4));
    BSM_CUDA(cudaMemsetAsync(*d, 0, kStatWords * 4, rt().stream));
    BSM_CUDA(cudaMemsetAsync(*d + kStatColMin, 0xFF, 4, rt().stream));
    return BSM_OK;
}

static int stats_finish(bsm_csr *a, uint32_t *d, bool check_cols)
{
    cudaStream_t sm = rt().stream;
    BSM_TRY(launch_csr_stats(a->row_ptr, a->col_idx, a->rows, a->nnz, a->row_offset, d, sm));
    std::vector<uint32_t> h(kStatWords);
    BSM_CUDA(cudaMemcpyAsync(h.data(), d, kStatWords * 4, cudaMemcpyDeviceToHost, sm));
    BSM_CUDA(cudaStreamSynchronize(sm));   // the only synchronisation of an upload
    if (h[kStatNarrowColBad]) return fail(BSM_ERR_OUT_OF_BOUNDS, "csr: a col_index is >= cols");
    if (h[kStatNarrowRowBad]) return fail(BSM_ERR_INVALID_ARGUMENT, "csr: row_index entry outside [0,nnz]");
    if (h[kStatBadRowPtr]) return fail(BSM_ERR_INVALID_ARGUMENT, "csr: row_index must start at 0, end at nnz and be non-decreasing");
    a->max_row_nnz = h[kStatMaxLen];
    a->col_min = a->nnz ? h[kStatColMin] : 0;
    a->col_max = a->nnz ? h[kStatColMax] : 0;
    if (check_cols && a->nnz && a->col_max >= a->cols) return fail(BSM_ERR_OUT_OF_BOUNDS, "csr: a col_index is >= cols");
    // dominant line length: the value most rows vote for, if at least half of all rows do
    a->row_stride = 0;
    if (a->rows >= 4096 && a->max_row_nnz >= 3 && a->max_row_nnz <= 64) {
        uint32_t best = 0, votes = 0;
        for (uint32_t s = kStrideMin; s <= kStrideMax; ++s)
            if (h[kStatHist + s] > votes) {
                votes = h[kStatHist + s];
                best = s;
            }
        if (best && (uint64_t)votes * 2 >= a->rows) a->row_stride = best;
    }
    return BSM_OK;
}

int compute_stats(bsm_csr *a, bool check_cols)
{
    PhaseScope ph(PH_A_STATS);
    uint32_t *d = nullptr;
    int st = stats_scratch_alloc(&d);
    if (st == BSM_OK) st = stats_finish(a, d, check_cols);
    tmp_free(d);
    return st;
}

template <typename T>
static int csr_upload_rows(int dtype, uint64_t rows_total, uint64_t cols, const T *v, const uint64_t *col_index,
                           const uint64_t *row_index, uint64_t row_begin, uint64_t row_end, bsm_csr **out)
{
    BSM_TRY(ensure_init());
    if (!out || !row_index || row_begin > row_end || row_end > rows_total)
        return fail(BSM_ERR_INVALID_ARGUMENT, "csr_upload: bad arguments");
    const uint64_t e0 = row_index[row_begin], e1 = row_index[row_end];
    if (e1 < e0) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_upload: row_index is not non-decreasing");
    const uint64_t nnz = e1 - e0, rows = row_end - row_begin;
    if (nnz && (!v || !col_index)) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_upload: null value/index arrays");
    bsm_csr *a = nullptr;
    BSM_TRY(alloc_csr(dtype, rows, cols, nnz, &a));
    a->row_offset = row_begin;
    cudaStream_t sm = rt().stream;
    uint64_t *stage_c = nullptr, *stage_r = nullptr;
    uint32_t *stats = nullptr;
    auto body = [&]() -> int {
        PhaseScope ph(PH_A_UPLOAD);
        BSM_TRY(stats_scratch_alloc(&stats));
        if (nnz) {
            BSM_TRY(tmp_alloc((void **)&stage_c, nnz * 8));
            BSM_CUDA(cudaMemcpyAsync(a->vals, v + e0, nnz * sizeof(T), cudaMemcpyHostToDevice, sm));
            // usize -> u32 narrowing and the column bound check happen on the device
            BSM_CUDA(cudaMemcpyAsync(stage_c, col_index + e0, nnz * 8, cudaMemcpyHostToDevice, sm));
            BSM_TRY(launch_narrow_u64(stage_c, a->col_idx, nnz, cols, 0, stats + kStatNarrowColBad, sm));
        }
        BSM_TRY(tmp_alloc((void **)&stage_r, (rows + 1) * 8));
        BSM_CUDA(cudaMemcpyAsync(stage_r, row_index + row_begin, (rows + 1) * 8, cudaMemcpyHostToDevice, sm));
        BSM_TRY(launch_narrow_u64(stage_r, a->row_ptr, rows + 1, nnz + 1, e0, stats + kStatNarrowRowBad, sm));
        return BSM_OK;
    };
    int st = body();
    if (st == BSM_OK) {
        PhaseScope ph(PH_A_STATS);
        st = stats_finish(a, stats, false);   // statistics + the narrowing flags: one readback, one synchronisation
    }
    tmp_free(stage_c);
    tmp_free(stage_r);
    tmp_free(stats);
    if (st != BSM_OK) {
        bsm_csr_free(a);
        return st;
    }
    *out = a;
    return BSM_OK;
}

template <typename T>
int csr_upload(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const T *v, const uint64_t *col_index,
                      const uint64_t *row_index, uint64_t row_index_len, bsm_csr **out)
{
    // a finalised reference Csr has row_index.len() == rows+1 and *row_index.last() == nnz
    // (src/sparse.rs:206-219, 162-164)
    if (row_index_len != rows + 1)
        return fail(BSM_ERR_NOT_FINALISED, "csr_upload: row_index must have rows+1 entries (call finalise() first)");
    if (!row_index) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_upload: null row_index");
    if (row_index[0] != 0 || row_index[rows] != nnz)
        return fail(BSM_ERR_INVALID_ARGUMENT, "csr_upload: row_index must start at 0 and end at nnz");
    return csr_upload_rows<T>(dtype, rows, cols, v, col_index, row_index, 0, rows, out);
}

template <typename T> int csr_download(const bsm_csr *a, int dtype, T *v, uint64_t *col_index, uint64_t *row_index)
{
    BSM_TRY(ensure_init());
    if (!a) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_download: null handle");
    if (a->dtype != dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "csr_download: dtype mismatch");
    cudaStream_t sm = rt().stream;
    uint64_t *stage = nullptr;
    const uint64_t stage_elems = std::max<uint64_t>(a->nnz, a->rows + 1);
    BSM_TRY(tmp_alloc((void **)&stage, stage_elems * 8));
    int st = [&]() -> int {
        if (a->nnz) {
            BSM_CUDA(cudaMemcpyAsync(v, a->vals, a->nnz * sizeof(T), cudaMemcpyDeviceToHost, sm));
            BSM_TRY(launch_widen_u32(a->col_idx, stage, a->nnz, sm));
            BSM_CUDA(cudaMemcpyAsync(col_index, stage, a->nnz * 8, cudaMemcpyDeviceToHost, sm));
        }
        BSM_TRY(launch_widen_u32(a->row_ptr, stage, a->rows + 1, sm));
        BSM_CUDA(cudaMemcpyAsync(row_index, stage, (a->rows + 1) * 8, cudaMemcpyDeviceToHost, sm));
        BSM_CUDA(cudaStreamSynchronize(sm));
        return BSM_OK;
    }();
    tmp_free(stage);
    return st;
}

int dense_alloc(int dtype, uint64_t rows, uint64_t cols, bsm_dense **out)
{
    BSM_TRY(ensure_init());
    if (!out) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_alloc: null out");
    if (dtype != BSM_F32 && dtype != BSM_F64) return fail(BSM_ERR_DTYPE_MISMATCH, "dense_alloc: dtype must be f32 or f64");
    if (rows >= 0xFFFFFFF0ull || cols >= 0xFFFFFFF0ull) return fail(BSM_ERR_INDEX_OVERFLOW, "dense dims must fit u32");
    bsm_dense *d = new bsm_dense();
    d->dtype = dtype;
    d->rows = rows;
    d->cols = cols;
    d->ld = default_ld(cols, dtype);
    const size_t bytes = (size_t)rows * d->ld * dtype_size(dtype);
    int st = dev_alloc(&d->data, bytes + 16, &d->pooled);
    if (st != BSM_OK) {
        delete d;
        return st;
    }
    if (d->ld != cols && bytes) cudaMemsetAsync(d->data, 0, bytes, rt().stream);   // defined padding columns
    *out = d;
    return BSM_OK;
}

constexpr uint64_t kColGroup = 32;   // columns staged per transpose step

template <typename T> int dense_upload(int dtype, uint64_t rows, uint64_t cols, const T *const *col_ptrs, bsm_dense **out)
{
    BSM_TRY(ensure_init());
    if (cols && !col_ptrs) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_upload: null column pointers");
    bsm_dense *d = nullptr;
    BSM_TRY(dense_alloc(dtype, rows, cols, &d));
    cudaStream_t sm = rt().stream;
    T *stage = nullptr;
    int st = [&]() -> int {
        if (rows == 0 || cols == 0) return BSM_OK;
        const uint64_t g = std::min<uint64_t>(kColGroup, cols);
        BSM_TRY(tmp_alloc((void **)&stage, g * rows * sizeof(T)));
        for (uint64_t c0 = 0; c0 < cols; c0 += g) {
            const uint64_t gc = std::min<uint64_t>(g, cols - c0);
            for (uint64_t c = 0; c < gc; ++c) {
                if (!col_ptrs[c0 + c]) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_upload: null column");
                BSM_CUDA(cudaMemcpyAsync(stage + c * rows, col_ptrs[c0 + c], rows * sizeof(T), cudaMemcpyHostToDevice, sm));
            }
            // column-major Vec<Vec<T>> (src/dense.rs:5-9) -> row-major device layout
            BSM_TRY(launch_transpose_cm2rm(dtype, stage, (T *)d->data + c0, rows, gc, d->ld, sm));
        }
        BSM_CUDA(cudaStreamSynchronize(sm));
        return BSM_OK;
    }();
    tmp_free(stage);
    if (st != BSM_OK) {
        bsm_dense_free(d);
        return st;
    }
    *out = d;
    return BSM_OK;
}

template <typename T> static int dense_download(const bsm_dense *d, int dtype, T *const *col_ptrs)
{
    BSM_TRY(ensure_init());
    if (!d) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_download: null handle");
    if (d->dtype != dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "dense_download: dtype mismatch");
    if (d->rows == 0 || d->cols == 0) return BSM_OK;
    if (!col_ptrs) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_download: null column pointers");
    cudaStream_t sm = rt().stream;
    T *stage = nullptr;
    const uint64_t g = std::min<uint64_t>(kColGroup, d->cols);
    BSM_TRY(tmp_alloc((void **)&stage, g * d->rows * sizeof(T)));
    int st = [&]() -> int {
        for (uint64_t c0 = 0; c0 < d->cols; c0 += g) {
            const uint64_t gc = std::min<uint64_t>(g, d->cols - c0);
            BSM_TRY(launch_transpose_rm2cm(dtype, (const T *)d->data + c0, stage, d->rows, gc, d->ld, sm));
            for (uint64_t c = 0; c < gc; ++c)
                BSM_CUDA(cudaMemcpyAsync(col_ptrs[c0 + c], stage + c * d->rows, d->rows * sizeof(T), cudaMemcpyDeviceToHost, sm));
            BSM_CUDA(cudaStreamSynchronize(sm));   // stage is reused by the next group
        }
        return BSM_OK;
    }();
    tmp_free(stage);
    return st;
}

// used by pipeline.cu
template int csr_upload<double>(int, uint64_t, uint64_t, uint64_t, const double *, const uint64_t *, const uint64_t *, uint64_t, bsm_csr **);
template int csr_upload<float>(int, uint64_t, uint64_t, uint64_t, const float *, const uint64_t *, const uint64_t *, uint64_t, bsm_csr **);
template int csr_download<double>(const bsm_csr *, int, double *, uint64_t *, uint64_t *);
template int csr_download<float>(const bsm_csr *, int, float *, uint64_t *, uint64_t *);
template int dense_upload<double>(int, uint64_t, uint64_t, const double *const *, bsm_dense **);
template int dense_upload<float>(int, uint64_t, uint64_t, const float *const *, bsm_dense **);

// dimensions and leading dimension of a caller-provided dense view must fit the kernels' 32-bit row arithmetic
static int check_dense_view(uint64_t rows, uint64_t cols, uint64_t ld, int dtype, const char *who)
{
    if (rows >= 0xFFFFFFF0ull || cols >= 0xFFFFFFF0ull || ld >= 0xFFFFFFF0ull) return fail(BSM_ERR_INDEX_OVERFLOW, std::string(who) + ": dense dims must fit u32");
    if (ld * dtype_size(dtype) >= (1ull << 32)) return fail(BSM_ERR_INDEX_OVERFLOW, std::string(who) + ": a row (ld * sizeof(T)) must be shorter than 4 GiB");
    return BSM_OK;
}

}  // namespace bsm

using namespace bsm;

extern "C" {

// ---- Csr --------------------------------------------------------------------------------------
int bsm_csr_upload_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v, const uint64_t *col_index,
                       const uint64_t *row_index, uint64_t row_index_len, bsm_csr **out)
{
    return csr_upload<double>(BSM_F64, rows, cols, nnz, v, col_index, row_index, row_index_len, out);
}
int bsm_csr_upload_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v, const uint64_t *col_index,
                       const uint64_t *row_index, uint64_t row_index_len, bsm_csr **out)
{
    return csr_upload<float>(BSM_F32, rows, cols, nnz, v, col_index, row_index, row_index_len, out);
}
int bsm_csr_upload_rows_f64(uint64_t rows, uint64_t cols, const double *v, const uint64_t *col_index,
                            const uint64_t *row_index, uint64_t row_begin, uint64_t row_end, bsm_csr **out)
{
    return csr_upload_rows<double>(BSM_F64, rows, cols, v, col_index, row_index, row_begin, row_end, out);
}
int bsm_csr_upload_rows_f32(uint64_t rows, uint64_t cols, const float *v, const uint64_t *col_index,
                            const uint64_t *row_index, uint64_t row_begin, uint64_t row_end, bsm_csr **out)
{
    return csr_upload_rows<float>(BSM_F32, rows, cols, v, col_index, row_index, row_begin, row_end, out);
}

int bsm_csr_from_device(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const void *d_vals, const uint32_t *d_col_idx,
                        const uint32_t *d_row_ptr, int copy, bsm_csr **out)
{
    BSM_TRY(ensure_init());
    if (!out || !d_row_ptr || (nnz && (!d_vals || !d_col_idx))) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_from_device: null argument");
    if (dtype != BSM_F32 && dtype != BSM_F64) return fail(BSM_ERR_DTYPE_MISMATCH, "csr_from_device: dtype must be f32 or f64");
    bsm_csr *a = nullptr;
    if (copy) {
        BSM_TRY(alloc_csr(dtype, rows, cols, nnz, &a));
        cudaStream_t sm = rt().stream;
        cudaError_t e = cudaSuccess;
        if (nnz) e = cudaMemcpyAsync(a->vals, d_vals, nnz * dtype_size(dtype), cudaMemcpyDeviceToDevice, sm);
        if (nnz && e == cudaSuccess) e = cudaMemcpyAsync(a->col_idx, d_col_idx, nnz * 4, cudaMemcpyDeviceToDevice, sm);
        if (e == cudaSuccess) e = cudaMemcpyAsync(a->row_ptr, d_row_ptr, (rows + 1) * 4, cudaMemcpyDeviceToDevice, sm);
        if (e != cudaSuccess) {
            bsm_csr_free(a);
            return fail(BSM_ERR_CUDA, std::string("csr_from_device: ") + cudaGetErrorString(e));
        }
    } else {
        if (rows >= 0xFFFFFFF0ull || cols >= 0xFFFFFFF0ull || nnz >= 0xFFFFFFF0ull)
            return fail(BSM_ERR_INDEX_OVERFLOW, "rows, cols and nnz must fit the device's u32 indices");
        if (((uintptr_t)d_vals | (uintptr_t)d_col_idx | (uintptr_t)d_row_ptr) & 15)
            return fail(BSM_ERR_INVALID_ARGUMENT, "csr_from_device: borrowed arrays must be 16-byte aligned");
        a = new bsm_csr();
        a->dtype = dtype;
        a->rows = rows;
        a->cols = cols;
        a->nnz = nnz;
        a->vals = const_cast<void *>(d_vals);
        a->col_idx = const_cast<uint32_t *>(d_col_idx);
        a->row_ptr = const_cast<uint32_t *>(d_row_ptr);
        a->owns = false;
    }
    // the same validation an upload gets: row_ptr[0] == 0, row_ptr[rows] == nnz, non-decreasing, every column < cols
    // (an out-of-range column would gather B rows out of bounds; a row_ptr end past nnz would over-read the TMA slices)
    int st = compute_stats(a, true);
    if (st != BSM_OK) {
        bsm_csr_free(a);
        return st;
    }
    *out = a;
    return BSM_OK;
}

int bsm_csr_free(bsm_csr *a)
{
    if (!a) return BSM_OK;
    if (a->owns) {
        dev_free(a->vals, a->pooled);
        dev_free(a->col_idx, a->pooled);
        dev_free(a->row_ptr, a->pooled);
    }
    dev_free(a->part_rows, a->cache_pooled);
    dev_free(a->carry_vals, a->cache_pooled);
    dev_free(a->long_rows, a->cache_pooled);
    delete a;
    return BSM_OK;
}

int bsm_csr_info(const bsm_csr *a, int *dtype, uint64_t *rows, uint64_t *cols, uint64_t *nnz, uint64_t *max_row_nnz)
{
    if (!a) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_info: null handle");
    if (dtype) *dtype = a->dtype;
    if (rows) *rows = a->rows;
    if (cols) *cols = a->cols;
    if (nnz) *nnz = a->nnz;
    if (max_row_nnz) *max_row_nnz = a->max_row_nnz;
    return BSM_OK;
}

int bsm_csr_stats(const bsm_csr *a, uint64_t *max_row_nnz, uint64_t *col_min, uint64_t *col_max, uint64_t *line_length)
{
    if (!a) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_stats: null handle");
    if (max_row_nnz) *max_row_nnz = a->max_row_nnz;
    if (col_min) *col_min = a->col_min;
    if (col_max) *col_max = a->col_max;
    if (line_length) *line_length = a->row_stride;
    return BSM_OK;
}

int bsm_csr_device_ptrs(const bsm_csr *a, const void **d_vals, const uint32_t **d_col_idx, const uint32_t **d_row_ptr)
{
    if (!a) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_device_ptrs: null handle");
    if (d_vals) *d_vals = a->vals;
    if (d_col_idx) *d_col_idx = a->col_idx;
    if (d_row_ptr) *d_row_ptr = a->row_ptr;
    return BSM_OK;
}

int bsm_csr_download_f64(const bsm_csr *a, double *v, uint64_t *col_index, uint64_t *row_index)
{
    return csr_download<double>(a, BSM_F64, v, col_index, row_index);
}
int bsm_csr_download_f32(const bsm_csr *a, float *v, uint64_t *col_index, uint64_t *row_index)
{
    return csr_download<float>(a, BSM_F32, v, col_index, row_index);
}

// ---- Dense ------------------------------------------------------------------------------------
int bsm_dense_upload_f64(uint64_t rows, uint64_t cols, const double *const *col_ptrs, bsm_dense **out)
{
    return dense_upload<double>(BSM_F64, rows, cols, col_ptrs, out);
}
int bsm_dense_upload_f32(uint64_t rows, uint64_t cols, const float *const *col_ptrs, bsm_dense **out)
{
    return dense_upload<float>(BSM_F32, rows, cols, col_ptrs, out);
}
int bsm_dense_alloc(int dtype, uint64_t rows, uint64_t cols, bsm_dense **out) { return dense_alloc(dtype, rows, cols, out); }

int bsm_dense_borrow(int dtype, uint64_t rows, uint64_t cols, void *d_rowmajor, uint64_t ld, bsm_dense **out)
{
    BSM_TRY(ensure_init());
    if (!out || (!d_rowmajor && rows && cols) || ld < cols) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_borrow: bad arguments");
    if (dtype != BSM_F32 && dtype != BSM_F64) return fail(BSM_ERR_DTYPE_MISMATCH, "dense_borrow: dtype must be f32 or f64");
    BSM_TRY(check_dense_view(rows, cols, ld, dtype, "dense_borrow"));
    bsm_dense *d = new bsm_dense();
    d->dtype = dtype;
    d->rows = rows;
    d->cols = cols;
    d->ld = ld ? ld : 1;
    d->data = d_rowmajor;
    d->owns = false;
    *out = d;
    return BSM_OK;
}

int bsm_dense_free(bsm_dense *d)
{
    if (!d) return BSM_OK;
    if (d->owns) dev_free(d->data, d->pooled);
    if (d->ipc && d->data) cudaIpcCloseMemHandle(d->data);
    delete d;
    return BSM_OK;
}

int bsm_dense_zero(bsm_dense *d)
{
    BSM_TRY(ensure_init());
    if (!d) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_zero: null handle");
    const size_t bytes = (size_t)d->rows * d->ld * dtype_size(d->dtype);
    if (bytes) BSM_CUDA(cudaMemsetAsync(d->data, 0, bytes, rt().stream));
    return BSM_OK;
}

int bsm_dense_info(const bsm_dense *d, int *dtype, uint64_t *rows, uint64_t *cols, uint64_t *ld, void **d_ptr)
{
    if (!d) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_info: null handle");
    if (dtype) *dtype = d->dtype;
    if (rows) *rows = d->rows;
    if (cols) *cols = d->cols;
    if (ld) *ld = d->ld;
    if (d_ptr) *d_ptr = d->data;
    return BSM_OK;
}

int bsm_dense_download_f64(const bsm_dense *d, double *const *col_ptrs) { return dense_download<double>(d, BSM_F64, col_ptrs); }
int bsm_dense_download_f32(const bsm_dense *d, float *const *col_ptrs) { return dense_download<float>(d, BSM_F32, col_ptrs); }

int bsm_dense_download_rowmajor(const bsm_dense *d, void *dst)
{
    BSM_TRY(ensure_init());
    if (!d || (!dst && d->rows && d->cols)) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_download_rowmajor: null argument");
    if (d->rows == 0 || d->cols == 0) return BSM_OK;
    const size_t s = dtype_size(d->dtype);
    BSM_CUDA(cudaMemcpy2DAsync(dst, d->cols * s, d->data, d->ld * s, d->cols * s, d->rows, cudaMemcpyDeviceToHost, rt().stream));
    BSM_CUDA(cudaStreamSynchronize(rt().stream));
    return BSM_OK;
}

int bsm_dense_upload_rowmajor(int dtype, uint64_t rows, uint64_t cols, const void *src, bsm_dense **out)
{
    bsm_dense *d = nullptr;
    BSM_TRY(dense_alloc(dtype, rows, cols, &d));
    if (rows && cols) {
        if (!src) {
            bsm_dense_free(d);
            return fail(BSM_ERR_INVALID_ARGUMENT, "dense_upload_rowmajor: null src");
        }
        const size_t s = dtype_size(dtype);
        cudaError_t e = cudaMemcpy2DAsync(d->data, d->ld * s, src, cols * s, cols * s, rows, cudaMemcpyHostToDevice, rt().stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(rt().stream);
        if (e != cudaSuccess) {
            bsm_dense_free(d);
            return fail(BSM_ERR_CUDA, std::string("dense_upload_rowmajor: ") + cudaGetErrorString(e));
        }
    }
    *out = d;
    return BSM_OK;
}

int bsm_dense_ipc_export(const bsm_dense *d, char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC memory handles are expected to be 64 bytes");
    BSM_TRY(ensure_init());
    if (!d || !handle) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_ipc_export: null argument");
    if (!d->owns || d->pooled || d->ipc) return fail(BSM_ERR_NOT_SUPPORTED, "dense_ipc_export: only buffers from bsm_dense_alloc / upload / gen can be exported");
    cudaIpcMemHandle_t h;
    BSM_CUDA(cudaIpcGetMemHandle(&h, d->data));
    memcpy(handle, &h, sizeof(h));
    return BSM_OK;
}

int bsm_dense_ipc_open(const char handle[64], int dtype, uint64_t rows, uint64_t cols, uint64_t ld, bsm_dense **out)
{
    BSM_TRY(ensure_init());
    if (!handle || !out || ld < cols) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_ipc_open: bad arguments");
    if (dtype != BSM_F32 && dtype != BSM_F64) return fail(BSM_ERR_DTYPE_MISMATCH, "dense_ipc_open: dtype must be f32 or f64");
    BSM_TRY(check_dense_view(rows, cols, ld, dtype, "dense_ipc_open"));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    BSM_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    bsm_dense *d = new bsm_dense();
    d->dtype = dtype;
    d->rows = rows;
    d->cols = cols;
    d->ld = ld ? ld : 1;
    d->data = p;
    d->owns = false;
    d->ipc = true;
    *out = d;
    return BSM_OK;
}

uint32_t bsm_line_length_of_row(const uint32_t *cols, uint32_t len, uint64_t diag) { return cols && len ? line_length_of_row(cols, len, diag) : 0u; }

}  // extern "C"
