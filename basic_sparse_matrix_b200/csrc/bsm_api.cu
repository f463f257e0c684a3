// bsm_api.cu — C ABI (include/bsm.h) over the sm_100a kernels: device handles, upload/download
// with format conversion, kernel dispatch heuristics, sharding helpers, generators.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bsm_common.cuh"
#include "kernels.h"

namespace bsm {

static Runtime g_rt;
static thread_local std::string g_err;
static thread_local bsm_launch_info g_info;
static std::atomic<uint64_t> g_launches{0};

Runtime &rt() { return g_rt; }
void set_error(const std::string &msg) { g_err = msg; }
int fail(int status, const std::string &msg)
{
    g_err = msg;
    return status;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int ensure_init()
{
    if (g_rt.device >= 0) return BSM_OK;
    return bsm_init(0);
}

// Two kinds of device memory:
//   * long-lived handles created by the public upload / alloc / generator calls: cudaMalloc;
//   * temporaries, and the handles the host-to-host convenience calls create and destroy inside one
//     call (PoolScope): the stream-ordered pool of the device, kept warm (release threshold = max), so
//     that a small multiplication does not pay a dozen cudaMalloc / cudaFree round trips.
static thread_local int g_pool_depth = 0;
struct PoolScope {
    PoolScope() { ++g_pool_depth; }
    ~PoolScope() { --g_pool_depth; }
};

static int tmp_alloc(void **p, size_t bytes)
{
    BSM_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, g_rt.stream));
    return BSM_OK;
}
static void tmp_free(void *p)
{
    if (p) cudaFreeAsync(p, g_rt.stream);
}
static int dev_alloc(void **p, size_t bytes, bool *pooled)
{
    *pooled = g_pool_depth > 0;
    if (*pooled) return tmp_alloc(p, bytes);
    BSM_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    return BSM_OK;
}
static void dev_free(void *p, bool pooled)
{
    if (!p) return;
    if (pooled)
        cudaFreeAsync(p, g_rt.stream);
    else
        cudaFree(p);
}

static inline uint64_t pad4(uint64_t n) { return (n + 3) / 4 * 4; }

// leading dimension of library-owned dense buffers: rows start 16-byte aligned whenever a row is
// at least 16 bytes, so lanes can use 128-bit loads
static uint64_t default_ld(uint64_t cols, int dtype)
{
    const uint64_t v = 16 / dtype_size(dtype);
    if (cols <= 1) return cols ? cols : 1;
    return round_up(cols, v);
}

static int alloc_csr(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, bsm_csr **out)
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
    cudaStream_t sm = g_rt.stream;
    cudaMemsetAsync((char *)a->vals + nnz * s, 0, (pad4(nnz) + 4 - nnz) * s, sm);
    cudaMemsetAsync(a->col_idx + nnz, 0, (pad4(nnz) + 4 - nnz) * 4, sm);
    cudaMemsetAsync(a->row_ptr + rows + 1, 0, (pad4(rows + 1) + 4 - rows - 1) * 4, sm);
    *out = a;
    return BSM_OK;
}

// Stencil-like matrices (Laplacians on a grid) touch B at fixed column offsets from the diagonal:
// +-1, +-nx, +-nx*ny. The smallest offset above 1 ("row stride") tells the vector kernel how many
// consecutive rows one warp should own so that the warps of a CTA sweep adjacent grid lines and
// share those B rows through L1. Sampled from three interior rows; 0 when the matrix is not of that
// shape. A performance hint only: any value is correct.
// Line length seen from one row: the smallest distance > 1 of a stored column from the diagonal (NOT from the row's
// median column: on a grid boundary row the median is a neighbour and the result is off by one — 4095 for a line of
// 4096). A box stencil (9- / 27-point) also stores the neighbours of its line neighbour — distances nx-1, nx, nx+1 —
// and yields nx. 0 when the row has no such column. Pure host arithmetic (bsm_line_length_of_row, CPU-tested).
static uint32_t line_length_of_row(const uint32_t *cols, uint32_t len, uint64_t diag)
{
    auto dist = [&](uint32_t i) { return cols[i] > diag ? cols[i] - diag : diag - cols[i]; };
    uint64_t stride = 0;
    for (uint32_t i = 0; i < len; ++i) {
        const uint64_t d = dist(i);
        if (d > 1 && (stride == 0 || d < stride)) stride = d;
    }
    if (stride == 0 || stride > 0xFFFFFFF0ull) return 0;
    bool plus1 = false, plus2 = false;
    for (uint32_t i = 0; i < len; ++i) {
        plus1 |= dist(i) == stride + 1;
        plus2 |= dist(i) == stride + 2;
    }
    return (uint32_t)(plus1 && plus2 ? stride + 1 : stride);
}

static int detect_row_stride(bsm_csr *a)
{
    a->row_stride = 0;
    if (a->rows < 4096 || a->max_row_nnz < 3 || a->max_row_nnz > 64) return BSM_OK;
    uint32_t found = 0;
    for (int q = 1; q <= 3; ++q) {
        const uint64_t r = a->rows / 4 * q;
        uint32_t rp[2];
        BSM_CUDA(cudaMemcpyAsync(rp, a->row_ptr + r, 8, cudaMemcpyDeviceToHost, g_rt.stream));
        BSM_CUDA(cudaStreamSynchronize(g_rt.stream));
        const uint32_t len = rp[1] - rp[0];
        if (len < 3 || len > 64) return BSM_OK;
        uint32_t cols[64];
        BSM_CUDA(cudaMemcpyAsync(cols, a->col_idx + rp[0], len * 4, cudaMemcpyDeviceToHost, g_rt.stream));
        BSM_CUDA(cudaStreamSynchronize(g_rt.stream));
        const uint32_t stride = line_length_of_row(cols, len, r + a->row_offset);
        if (stride < 16 || stride > 16384 || (found && stride != found)) return BSM_OK;
        found = stride;
    }
    a->row_stride = found;
    return BSM_OK;
}

static int compute_stats(bsm_csr *a)
{
    uint32_t *d = nullptr;
    BSM_CUDA(cudaMallocAsync(&d, 16, g_rt.stream));
    const uint32_t init[4] = {0u, 0u, 0xFFFFFFFFu, 0u};   // max row length, bad flag, min column, max column
    BSM_CUDA(cudaMemcpyAsync(d, init, 16, cudaMemcpyHostToDevice, g_rt.stream));
    BSM_TRY(launch_row_stats(a->row_ptr, a->rows, d, d + 1, g_rt.stream));
    BSM_TRY(launch_col_range(a->col_idx, a->nnz, d + 2, g_rt.stream));
    uint32_t h[4] = {0, 0, 0, 0};
    BSM_CUDA(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, g_rt.stream));
    BSM_CUDA(cudaFreeAsync(d, g_rt.stream));
    BSM_CUDA(cudaStreamSynchronize(g_rt.stream));
    if (h[1]) return fail(BSM_ERR_INVALID_ARGUMENT, "row_index is not non-decreasing");
    a->max_row_nnz = h[0];
    a->col_min = a->nnz ? h[2] : 0;
    a->col_max = a->nnz ? h[3] : 0;
    return detect_row_stride(a);
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
    cudaStream_t sm = g_rt.stream;
    uint64_t *stage = nullptr;
    uint32_t *flags = nullptr;
    const uint64_t stage_elems = std::max<uint64_t>(nnz, rows + 1);
    int st = BSM_OK;
    auto body = [&]() -> int {
        BSM_TRY(tmp_alloc((void **)&stage, stage_elems * 8));
        BSM_TRY(tmp_alloc((void **)&flags, 8));
        BSM_CUDA(cudaMemsetAsync(flags, 0, 8, sm));
        if (nnz) {
            BSM_CUDA(cudaMemcpyAsync(a->vals, v + e0, nnz * sizeof(T), cudaMemcpyHostToDevice, sm));
            // usize -> u32 narrowing and the column bound check happen on the device
            BSM_CUDA(cudaMemcpyAsync(stage, col_index + e0, nnz * 8, cudaMemcpyHostToDevice, sm));
            BSM_TRY(launch_narrow_u64(stage, a->col_idx, nnz, cols, 0, flags, sm));
        }
        BSM_CUDA(cudaMemcpyAsync(stage, row_index + row_begin, (rows + 1) * 8, cudaMemcpyHostToDevice, sm));
        BSM_TRY(launch_narrow_u64(stage, a->row_ptr, rows + 1, nnz + 1, e0, flags + 1, sm));
        uint32_t h[2] = {0, 0};
        BSM_CUDA(cudaMemcpyAsync(h, flags, 8, cudaMemcpyDeviceToHost, sm));
        BSM_CUDA(cudaStreamSynchronize(sm));
        if (h[0]) return fail(BSM_ERR_OUT_OF_BOUNDS, "csr_upload: a col_index is >= cols");
        if (h[1]) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_upload: row_index entry outside [0,nnz]");
        return compute_stats(a);
    };
    st = body();
    tmp_free(stage);
    tmp_free(flags);
    if (st != BSM_OK) {
        bsm_csr_free(a);
        return st;
    }
    *out = a;
    return BSM_OK;
}

template <typename T>
static int csr_upload(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const T *v, const uint64_t *col_index,
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

template <typename T> static int csr_download(const bsm_csr *a, int dtype, T *v, uint64_t *col_index, uint64_t *row_index)
{
    BSM_TRY(ensure_init());
    if (!a) return fail(BSM_ERR_INVALID_ARGUMENT, "csr_download: null handle");
    if (a->dtype != dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "csr_download: dtype mismatch");
    cudaStream_t sm = g_rt.stream;
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

static int dense_alloc(int dtype, uint64_t rows, uint64_t cols, bsm_dense **out)
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
    if (d->ld != cols && bytes) cudaMemsetAsync(d->data, 0, bytes, g_rt.stream);   // defined padding columns
    *out = d;
    return BSM_OK;
}

constexpr uint64_t kColGroup = 32;   // columns staged per transpose step

template <typename T> static int dense_upload(int dtype, uint64_t rows, uint64_t cols, const T *const *col_ptrs, bsm_dense **out)
{
    BSM_TRY(ensure_init());
    if (cols && !col_ptrs) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_upload: null column pointers");
    bsm_dense *d = nullptr;
    BSM_TRY(dense_alloc(dtype, rows, cols, &d));
    cudaStream_t sm = g_rt.stream;
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
    cudaStream_t sm = g_rt.stream;
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

// ------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------
static int pow2_ceil(int x)
{
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// lane shape for `n` columns starting at byte-aligned pointers
static Shape pick_shape(uint32_t n, uint64_t ldb, uint64_t ldc, uint64_t col0, const void *b, const void *c, size_t s,
                        bool prefer_wide, uint64_t extra_ld = 0)
{
    Shape sh;
    int V = (int)(16 / s);
    auto ok = [&](int v) {
        const uint64_t bytes = (uint64_t)v * s;
        return n % v == 0 && ldb % v == 0 && ldc % v == 0 && col0 % v == 0 && (extra_ld % v == 0) &&
               ((uintptr_t)b % bytes == 0) && ((uintptr_t)c % bytes == 0);
    };
    while (V > 1 && !ok(V)) V /= 2;
    int L = (int)(n / V);
    if (prefer_wide)
        while (L < 32 && V > 1) {
            V /= 2;
            L = (int)(n / V);
        }
    sh.V = V;
    if (L >= 32) {
        sh.G = 32;
        const int nt = (L + 31) / 32;
        sh.NT = nt <= 1 ? 1 : (nt <= 2 ? 2 : 4);
    } else {
        sh.G = pow2_ceil(L);
        sh.NT = 1;
    }
    return sh;
}

// Columns one pass can take starting at col0 when `want` remain (<= the tile): the lane shape holds at most
// G * V * NT columns, and alignment can force vectors narrower than 16 bytes — 129 f32 columns are 129 one-element
// lanes, more than the 128 that four register tiles hold. Then the pass is cut to a width the widest vectors divide
// (128 of the 129; the last column goes to the next pass), or failing that to what the narrow vectors hold.
template <typename ShapeOf> static uint32_t fit_pass_width(uint32_t want, int vmax, ShapeOf &&shape_of)
{
    auto holds = [](const Shape &x) { return (uint32_t)(x.G * x.V * x.NT); };
    const Shape sh = shape_of(want);
    if (holds(sh) >= want) return want;
    const uint32_t even = want / (uint32_t)vmax * (uint32_t)vmax;
    if (even >= (uint32_t)vmax && even < want && holds(shape_of(even)) >= even) return even;
    return holds(sh);
}

// Geometry of the vector kernel for one column pass (see spmm_rows.cu).
// Scatter variant of the vector kernel: C is this rank's FULL result buffer, the rank's rows start at
// row_offset, and every row is also stored to `n_peers` further full buffers (peer GPUs over NVLink).
struct ScatterTargets {
    int n_peers = 0;
    void *peer_data[7] = {};
    uint64_t row_offset = 0;
};

// ---- launch plan of the vector kernel for one column pass ------------------------------------------------
// Every default below comes from a same-box A/B sweep kept in profiles/ (tools/sweep.py); bsm_tuning overrides each.

// What the planner knows about the matrix and the device — no pointers, no CUDA calls, so the same code plans a
// launch for bsm_spmm and answers bsm_plan_vector (a dry run the CPU test-suite uses to pin the heuristics).
struct MatrixFacts {
    int dtype;
    uint64_t rows, nnz, max_row_nnz;
    uint32_t row_stride;   // line length of a stencil-like matrix (0 = none)
    double mean() const { return rows ? (double)nnz / (double)rows : 0.0; }
};
struct DeviceFacts {
    int sm_count;
    size_t smem_max;       // dynamic shared memory one CTA may use
};
static MatrixFacts facts_of(const bsm_csr *a) { return MatrixFacts{a->dtype, a->rows, a->nnz, a->max_row_nnz, a->row_stride}; }
static double mean_row_nnz(const bsm_csr *a) { return facts_of(a).mean(); }

// Grouped lanes: fewer lanes per row than the 128-bit loads need -> 2 or 4 register tiles per lane and 32/G rows side
// by side, every lane group walking its own flat entry stream over a run of consecutive rows (one LDS of the staged A
// stream then feeds 32/G rows). Exists for full-width 128-bit shapes only. Returns true when `sh` was regrouped;
// `by_default` says the choice was the heuristic's, not the caller's.
//   defaults on regular rows (uneven rows stay on the warp-per-row stream, which balances them inside the warp;
//   3-D Laplacian, profiles/r1_sweep{u,v,w,y,z,aa}_l3d_*.jsonl):
//     one 128-bit tile per lane, 512-byte rows (x64 f64): 8 lanes x 4 tiles            4.54 -> 3.92 ms
//     256-byte rows (x32 f64, x64 f32), short rows: 16 -> 8 lanes x 2 tiles            4.28 -> 2.22 ms
//     128-byte rows (x16 f64), short rows: 8 -> 4 lanes x 2 tiles                      2.70 -> 1.37 ms
//   (the row-by-row walk of narrow shapes drains its gather window at every row end; rows of ~65 entries, the band
//   matrix x32 f32, are still faster row by row: 0.45 vs 0.72 ms)
static bool regroup_lanes(const MatrixFacts &m, const bsm_tuning &tn, uint32_t n, size_t s, bool allowed, Shape &sh, bool &by_default)
{
    const double mean = m.mean();
    int want_g = tn.lanes_per_row;
    by_default = false;
    if (want_g == 0 && sh.NT == 1 && tn.reg_flavour <= 0 && tn.warps_per_cta <= 0 && tn.prefer_wide_rows == 0 &&
        (double)m.max_row_nnz <= 4.0 * mean + 8.0) {
        if (sh.G == 32) want_g = 8;
        else if ((sh.G == 16 || sh.G == 8) && mean <= 32.0) want_g = sh.G / 2;
        by_default = want_g > 0;
    }
    if (!allowed || want_g <= 0 || want_g >= sh.G || sh.V * (int)s != 16 || n != (uint32_t)(sh.V * sh.G * sh.NT)) return false;
    const int nt = (int)(n / (uint32_t)(sh.V * want_g));
    if (!(want_g == 16 || want_g == 8 || want_g == 4) || !(nt == 2 || nt == 4) || n != (uint32_t)(sh.V * want_g * nt)) return false;
    sh.G = want_g;
    sh.NT = nt;
    return true;
}

// Register-budget flavour (index into the table of spmm_rows_inst.cuh; bsm_tuning.reg_flavour is this + 1).
//   full-width G == 32: several tiles per lane -> 3 CTAs x 8 warps per SM (4); one tile per lane -> one CTA of 24 warps,
//   window of 10 gathers (7); scalar A-stream reads in both (LDS.128 reads measured 5-8 % slower);
//   grouped lanes: one CTA of 24 warps (6) for >= 8 lanes by default, else 3 x 8 warps (4);
//   narrow one-tile shapes: 0, or 4 = scalar A-stream reads (the default for a row per lane);
//   the scatter variant exists for the default flavours only.
static int pick_row_flavour(const bsm_tuning &tn, const Shape &sh, bool wide_full, bool grouped, bool grouped_by_default, bool multi)
{
    const bool user_nw = tn.warps_per_cta > 0;
    int flavour = tn.reg_flavour > 0 ? std::min(tn.reg_flavour, 8) - 1 : (wide_full && !user_nw ? (sh.NT >= 2 ? 4 : 7) : (sh.G == 1 ? 4 : 0));
    if (!wide_full && !(sh.G < 32 && flavour == 4)) flavour = 0;
    if (grouped) flavour = (tn.reg_flavour == 7 || tn.reg_flavour == 8 || (grouped_by_default && tn.reg_flavour <= 0 && sh.G >= 8)) ? 6 : 4;
    if (flavour == 3) flavour = 2;                       // retired flavour
    if (flavour == 7 && sh.NT >= 2) flavour = 4;         // the deep window exists for one tile per lane only
    if (multi) flavour = wide_full ? (sh.NT >= 2 ? 4 : 2) : 0;
    return flavour;
}

// Rows per warp inside a super-batch: the dominant row stride of a stencil-like matrix (so the warps of a CTA sweep
// adjacent grid lines), else one slice (a few slices for narrow shapes on large matrices: measured on band x 32).
static uint32_t pick_rows_per_warp(const MatrixFacts &m, const DeviceFacts &dev, const bsm_tuning &tn, const Shape &sh, uint32_t R, int nw, int resident)
{
    uint32_t P = R;
    if (tn.rows_per_warp <= 0 && sh.G < 32 && m.rows / ((uint64_t)nw * 4 * R) >= 4ull * dev.sm_count) P = 4 * R;
    if (tn.rows_per_warp > 0) {
        P = (uint32_t)tn.rows_per_warp;
    } else if (m.row_stride >= 2 * R) {
        // P = stride / m keeps warps w and w+m on adjacent lines. Among stride, stride/2, stride/4, ... pick the one that
        // wastes least to wave quantisation (rounds x rows per warp per round); a larger P wins unless a smaller one
        // saves more than 10 % (measured on 1/8 and 1/4 row blocks: profiles/r1_sweepk_l3d_n128_s8.jsonl — locality
        // beats balance). Too few rows for even one round per SM: plain slices.
        const uint64_t grid_est = (uint64_t)dev.sm_count * resident;
        double best_cost = 0.0;
        uint32_t best_p = 0;
        for (uint32_t cand = m.row_stride; cand >= 2 * R; cand /= 2) {
            const uint64_t supers = (m.rows + (uint64_t)nw * cand - 1) / ((uint64_t)nw * cand);
            const double cost = (double)((supers + grid_est - 1) / grid_est) * cand;
            if (best_p == 0 || cost < 0.90 * best_cost) {
                best_cost = cost;
                best_p = cand;
            }
            if (cand % 2) break;
        }
        P = best_p ? best_p : R;
        if (m.rows / ((uint64_t)nw * P) < (uint64_t)dev.sm_count) P = R;
    }
    // flat-stream shapes take any P >= R (the last slice of a line may be short; the row_ptr windows are realigned in
    // the kernel); the row-by-row narrow shapes keep whole slices
    if (sh.G == 32 || sh.NT > 1) return std::max(R, P);
    return std::max(R, (P + R - 1) / R * R);
}

struct PassAlign {       // what pick_shape needs to know about the operands of one column pass
    uint64_t ldb, ldc, col0;
    const void *b, *c;
};
struct VectorPlan {
    Shape sh;
    bool grouped = false;
    int flavour = 0, nw = 0;
    uint32_t R = 0, P = 0, stages = 0, cap = 0, num_super = 0;
    size_t smem = 0;
    int resident = 1;    // CTAs per SM the flavour targets
};

// Launch plan of the vector kernel for one pass of n columns. Pure host arithmetic.
static int plan_vector_pass(const MatrixFacts &m, const DeviceFacts &dev, const bsm_tuning &tn, uint32_t n, const PassAlign &al, bool scatter,
                            bool multi, VectorPlan *out)
{
    const size_t s = dtype_size(m.dtype);
    const double mean = m.mean();
    const bool user_R = tn.rows_per_slice > 0, user_nw = tn.warps_per_cta > 0, user_stages = tn.stages > 0;
    const size_t smem_max = dev.smem_max;
    RowParams p{};   // geometry fields only (row_kernel_smem_bytes reads cap, R, stages)
    // second attempt = without grouped lanes, when their (always staged) slices do not fit shared memory
    for (bool allow_grouped = true;; allow_grouped = false) {
        Shape sh = pick_shape(n, al.ldb, al.ldc, al.col0, al.b, al.c, s, tn.prefer_wide_rows != 0);
        bool grouped_by_default = false;
        const bool grouped = regroup_lanes(m, tn, n, s, allow_grouped && !scatter, sh, grouped_by_default);
        const bool wide_full = sh.G == 32 && n == (uint32_t)(sh.V * sh.G * sh.NT);
        const uint32_t rpp = 32u / (uint32_t)sh.G;   // rows side by side in one warp
        const uint32_t rq = std::max(4u, rpp);       // slice granularity (rpp is a power of two)

        // rows per TMA slice: ~128 entries per bulk copy (~224 with one register tile per lane, r1_sweepi_*, and for
        // 8 lanes x 2 tiles); narrow shapes want several row passes per slice to amortise the slice bookkeeping
        uint32_t R;
        if (user_R) {
            R = (uint32_t)tn.rows_per_slice;
        } else {
            const double target = ((sh.G == 32 && sh.NT == 1) || (grouped && grouped_by_default && sh.NT == 2)) ? 224.0 : 128.0;
            R = (uint32_t)std::min<double>(256.0, std::max(1.0, target / std::max(1.0, mean)));
            if (sh.G < 32) R = std::max(R, 4u * rpp);
        }
        R = std::max(rq, R / rq * rq);
        // a stencil-like matrix: the slice must divide the line length, or the rows per warp (a multiple of the
        // slice) stop matching the lines and the L1 sharing between the warps of a CTA is lost (measured on a
        // 5-entry-per-row stencil, line 256: 24-row slices -> P = 264: 12.9 ms; 16-row slices -> P = 256: see
        // profiles/r1_probe_near_diag.jsonl)
        // (a divisor down to half the target; failing that the last slice of every line is short — a line of 100
        // keeps 16-row slices, 6 x 16 + 4, rather than dropping to 4-row slices)
        if (!user_R && m.row_stride >= 2 * rq && m.row_stride % R) {
            uint32_t r2 = R;
            while (r2 > rq && 2 * r2 > R && m.row_stride % r2) r2 -= rq;
            if (m.row_stride % r2 == 0) R = r2;
        }

        int flavour = pick_row_flavour(tn, sh, wide_full, grouped, grouped_by_default, multi);
        // warps per CTA: what the flavour was compiled for; fewer on small matrices, so that no SM idles behind a
        // handful of fat super-batches
        const bool big_cta = (wide_full && (flavour == 5 || flavour == 6 || flavour == 7)) || (grouped && flavour == 6);
        const int max_warps = big_cta ? 24 : ((flavour >= 2 && (wide_full || grouped)) ? 8 : 16);
        int nw = user_nw ? std::min(tn.warps_per_cta, 24) : (big_cta ? 24 : 16);
        if (!user_nw) {
            const uint64_t rows_per_warp_min = big_cta ? R : rq;
            const int nw_floor = big_cta ? 3 : 2;
            while (nw > nw_floor && (uint64_t)nw * rows_per_warp_min * (uint64_t)dev.sm_count > m.rows) nw /= 2;
        }
        nw = std::min(nw, max_warps);

        // The stage must hold the entries of ANY R consecutive rows (+3 for the 16-byte aligned start, + slack: the
        // vectorised A-stream reads run up to two gather windows past the slice). Shrink, in this order, the ring
        // depth, the slice and the CTA until the rings fit: first under a soft limit that leaves most of the 228 KB
        // to L1 (where wide B rows live), then under the hardware limit. If even the smallest slice cannot be
        // staged, col_idx / values are read from global memory instead (unstaged variant).
        const int resident = ((wide_full || grouped) && (flavour == 2 || flavour == 4)) ? 3 : 1;   // CTAs per SM the flavour targets
        const size_t smem_soft = (size_t)n * s <= 64 ? smem_max : std::min<size_t>(smem_max, (160 * 1024) / resident);
        const uint32_t window = flavour == 7 ? 10u : (sh.NT >= 4 ? 2u : (sh.NT == 2 ? 4u : 8u)) * (flavour == 1 ? 2u : 1u);   // gathers in flight
        const uint32_t r_floor = sh.G < 32 ? std::max(rq, 2u * rpp) : rq;
        p.R = R;
        p.stages = user_stages ? (uint32_t)std::min(tn.stages, 8) : (grouped && grouped_by_default && sh.G >= 8 ? 2u : 3u);
        auto smem_now = [&]() {
            p.cap = (uint32_t)pad4((uint64_t)p.R * m.max_row_nnz + 3) + 2 * window + 4;
            return row_kernel_smem_bytes(m.dtype, p, nw);
        };
        size_t smem = smem_now();
        while (smem > smem_soft && p.stages > 2 && !user_stages) { --p.stages; smem = smem_now(); }
        while (smem > smem_soft && p.R > r_floor && !user_R) { p.R = std::max(r_floor, p.R / 2 / rq * rq); smem = smem_now(); }
        while (smem > smem_soft && nw > 8 && !user_nw) { nw /= 2; smem = smem_now(); }
        while (smem > smem_max && p.stages > 1) { --p.stages; smem = smem_now(); }
        while (smem > smem_max && p.R > rq && !user_R) { p.R = std::max(rq, p.R / 2 / rq * rq); smem = smem_now(); }
        while (smem > smem_max && nw > 2 && !user_nw) { nw /= 2; smem = smem_now(); }
        if (smem > smem_max && grouped) continue;   // the grouped shapes have no unstaged variant: a warp per row instead
        if (smem > smem_max) {                      // rows too long to stage: unstaged variant (row_ptr windows only)
            flavour = -1;
            p.cap = 0;
            p.stages = user_stages ? (uint32_t)std::min(tn.stages, 8) : 3u;
            smem = row_kernel_smem_bytes(m.dtype, p, nw);
        }
        if (smem > smem_max) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_vector: slice ring does not fit shared memory");

        p.P = pick_rows_per_warp(m, dev, tn, sh, p.R, nw, resident);
        const uint64_t S = (uint64_t)nw * p.P;
        out->sh = sh;
        out->grouped = grouped;
        out->flavour = flavour;
        out->nw = nw;
        out->R = p.R;
        out->P = p.P;
        out->stages = p.stages;
        out->cap = p.cap;
        out->num_super = (uint32_t)((m.rows + S - 1) / S);
        out->smem = smem;
        out->resident = resident;
        return BSM_OK;
    }
}

static int spmm_vector(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning &tn, uint32_t flags,
                       const ScatterTargets *scatter = nullptr)
{
    const size_t s = dtype_size(a->dtype);
    const uint32_t n_total = (uint32_t)b->cols;
    const int vmax = (int)(16 / s);
    uint32_t tile = tn.col_tile > 0 ? (uint32_t)tn.col_tile : n_total;
    tile = std::min<uint32_t>(tile, 32u * vmax * 4u);   // widest shape one pass can hold in registers
    if (tile < n_total && tile % vmax) tile = std::max<uint32_t>(vmax, tile / vmax * vmax);
    const bool multi = scatter && scatter->n_peers > 0;
    const MatrixFacts m = facts_of(a);
    const DeviceFacts dev{g_rt.sm_count, (size_t)g_rt.max_smem_optin - 1024};
    int passes = 0;
    g_info = bsm_launch_info();
    g_info.algo = BSM_ALGO_VECTOR;
    g_info.col_tile = (int)tile;
    for (uint32_t col0 = 0, n = 0; col0 < n_total; col0 += n, ++passes) {
        n = fit_pass_width(std::min(tile, n_total - col0), vmax,
                           [&](uint32_t w) { return pick_shape(w, b->ld, c->ld, col0, b->data, c->data, s, tn.prefer_wide_rows != 0); });
        const size_t c_off = ((scatter ? (size_t)scatter->row_offset * c->ld : 0) + (size_t)col0) * s;
        RowParams p{};
        p.row_ptr = a->row_ptr;
        p.col_idx = a->col_idx;
        p.vals = a->vals;
        p.B = (const char *)b->data + (size_t)col0 * s;
        p.C = (char *)c->data + c_off;
        p.rows = (uint32_t)a->rows;
        p.n = n;
        p.ldb = (uint32_t)b->ld;
        p.ldc = (uint32_t)c->ld;
        p.flags = flags;
        if (multi) {
            p.n_peers = (uint32_t)scatter->n_peers;
            for (int d = 0; d < scatter->n_peers; ++d) p.peers[d] = (char *)scatter->peer_data[d] + c_off;
        }
        VectorPlan plan;
        BSM_TRY(plan_vector_pass(m, dev, tn, n, PassAlign{b->ld, c->ld, col0, b->data, c->data}, scatter != nullptr, multi, &plan));
        const Shape sh = plan.sh;
        const int flavour = plan.flavour, nw = plan.nw;
        const size_t smem = plan.smem;
        p.R = plan.R;
        p.P = plan.P;
        p.stages = plan.stages;
        p.cap = plan.cap;
        p.num_super = plan.num_super;
        {
            const int block = nw * 32;
            int occ = 0;
            BSM_TRY(row_kernel_occupancy(a->dtype, sh, n, flavour, multi, block, smem, &occ));
            if (occ < 1) return fail(BSM_ERR_CUDA, "spmm_vector: kernel does not fit on an SM");
            const int ctas = tn.ctas_per_sm > 0 ? std::min(tn.ctas_per_sm, occ) : std::min(occ, 4);
            const int grid = (int)std::min<uint64_t>(p.num_super, (uint64_t)g_rt.sm_count * ctas);
            if (grid > 0) BSM_TRY(launch_spmm_rows(a->dtype, sh, p, flavour, multi, grid, block, smem, ctas, g_rt.stream));
            g_info.kernels += grid > 0;
            g_info.vec_elems = sh.V;
            g_info.lanes_per_row = sh.G;
            g_info.reg_tiles = sh.NT;
            g_info.grid = grid;
            g_info.block = block;
            g_info.smem_bytes = (int)smem;
            g_info.rows_per_slice = (int)p.R;
            g_info.rows_per_warp = (int)p.P;
            g_info.reg_flavour = flavour + 1;
            g_info.stages = (int)p.stages;
            g_info.capacity = (int)p.cap;
        }
    }
    g_info.passes = passes;
    return BSM_OK;
}

// merge-path caches of a handle live where the handle's arrays live (pool or cudaMalloc); all of one kind
static int cache_alloc(bsm_csr *a, void **p, size_t bytes)
{
    if (!a->part_rows && !a->carry_vals && !a->long_rows) a->cache_pooled = a->pooled && a->owns;
    if (a->cache_pooled) return tmp_alloc(p, bytes);
    BSM_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    return BSM_OK;
}

static int ensure_partition(bsm_csr *a, uint32_t items, uint32_t num_chunks)
{
    if (a->part_rows && a->part_items == (int)items && a->part_chunks == num_chunks) return BSM_OK;
    dev_free(a->part_rows, a->cache_pooled);
    a->part_rows = nullptr;
    BSM_TRY(cache_alloc(a, (void **)&a->part_rows, ((size_t)num_chunks + 1) * 4));
    BSM_TRY(launch_merge_partition(a->row_ptr, (uint32_t)a->rows, (uint32_t)a->nnz, items, num_chunks, a->part_rows, g_rt.stream));
    g_info.kernels += 1;
    a->part_items = (int)items;
    a->part_chunks = num_chunks;
    return BSM_OK;
}

static int spmm_merge(const bsm_csr *a_const, const bsm_dense *b, bsm_dense *c, const bsm_tuning &tn, uint32_t flags)
{
    bsm_csr *a = const_cast<bsm_csr *>(a_const);   // partition / carry caches live in the handle
    const size_t s = dtype_size(a->dtype);
    const uint32_t n_total = (uint32_t)b->cols;
    const int vmax = (int)(16 / s);
    uint32_t tile = tn.col_tile > 0 ? (uint32_t)tn.col_tile : n_total;
    tile = std::min<uint32_t>(tile, 32u * vmax * 4u);
    if (tile < n_total && tile % vmax) tile = std::max<uint32_t>(vmax, tile / vmax * vmax);
    const uint64_t total = a->rows + a->nnz;
    g_info = bsm_launch_info();
    g_info.algo = BSM_ALGO_MERGE;
    int passes = 0;
    for (uint32_t col0 = 0, n = 0; col0 < n_total; col0 += n, ++passes) {
        n = fit_pass_width(std::min(tile, n_total - col0), vmax, [&](uint32_t w) {
            return pick_shape(w, b->ld, c->ld, col0, b->data, c->data, s, tn.prefer_wide_rows >= 0, round_up(w, vmax));
        });
        const uint64_t ldcar = round_up(n, vmax);
        Shape sh = pick_shape(n, b->ld, c->ld, col0, b->data, c->data, s, tn.prefer_wide_rows >= 0, ldcar);
        // fewer lanes per chunk, 2 or 4 register tiles per lane (128-bit lanes, full-width shapes): one LDS.128 of the
        // staged A stream then feeds 32/G chunks. Defaults from the same-box A/B on R-MAT (profiles/r1_sweepx_rmat_*):
        // 512-byte rows 16 lanes x 2 tiles, 192 items (3.08 -> 3.01 ms); 256-byte rows 8 lanes x 2 tiles, 160 items
        // (2.08 -> 1.84 ms, r1_sweepy_rmat_f32)
        int want_g = tn.lanes_per_row;
        uint32_t auto_items = 0;
        if (want_g == 0 && tn.prefer_wide_rows == 0 && tn.merge_items <= 0 && tn.warps_per_cta <= 0) {
            if ((size_t)n * s == 512) { want_g = 16; auto_items = 192; }
            if ((size_t)n * s == 256) { want_g = 8; auto_items = 160; }
        }
        bool grouped = false;
        if (want_g > 0) {
            const Shape sv = pick_shape(n, b->ld, c->ld, col0, b->data, c->data, s, false, ldcar);
            const int g = want_g, nt = sv.V * (int)s == 16 ? (int)(n / (uint32_t)(sv.V * g)) : 0;
            if ((g == 16 || g == 8 || g == 4) && (nt == 2 || nt == 4) && n == (uint32_t)(sv.V * g * nt)) {
                sh.V = sv.V;
                sh.G = g;
                sh.NT = nt;
                grouped = true;
            }
        }
        const int nw = tn.warps_per_cta > 0 ? std::min(tn.warps_per_cta, 8) : 8;
        const int block = nw * 32;
        const uint32_t groups = (uint32_t)nw * (32u / sh.G);
        uint32_t items = tn.merge_items > 0 ? (uint32_t)tn.merge_items : (grouped && auto_items ? auto_items : (sh.G == 32 ? 384u : std::max(16u, 2048u / groups)));
        items = (uint32_t)round_up(items, 4);
        if (total + items >= 0xFFFFFFF0ull) return fail(BSM_ERR_INDEX_OVERFLOW, "spmm_merge: rows+nnz must fit u32");
        const uint32_t num_chunks = (uint32_t)((total + items - 1) / items);
        if (num_chunks == 0) continue;
        BSM_TRY(ensure_partition(a, items, num_chunks));
        const size_t need_vals = (size_t)num_chunks * ldcar * s;
        if (a->carry_vals_bytes < need_vals) {
            dev_free(a->carry_vals, a->cache_pooled);
            a->carry_vals = nullptr;
            a->carry_vals_bytes = 0;
            BSM_TRY(cache_alloc(a, &a->carry_vals, need_vals));
            a->carry_vals_bytes = need_vals;
        }
        // rows whose run of carries exceeds 64 chunks need more than 64*items entries: at most this many
        const uint64_t long_cap = a->nnz / (64ull * items) + 1;
        if (a->long_rows_cap < long_cap) {
            dev_free(a->long_rows, a->cache_pooled);
            a->long_rows = nullptr;
            a->long_rows_cap = 0;
            BSM_TRY(cache_alloc(a, (void **)&a->long_rows, (2 * long_cap + 1) * 4));
            a->long_rows_cap = long_cap;
        }
        MergeParams p{};
        p.row_ptr = a->row_ptr;
        p.col_idx = a->col_idx;
        p.vals = a->vals;
        p.B = (const char *)b->data + (size_t)col0 * s;
        p.C = (char *)c->data + (size_t)col0 * s;
        p.part_rows = a->part_rows;
        p.carry_vals = a->carry_vals;
        p.long_rows = a->long_rows;
        p.long_count = a->long_rows + 2 * a->long_rows_cap;
        p.long_cap = (uint32_t)a->long_rows_cap;
        p.rows = (uint32_t)a->rows;
        p.nnz = (uint32_t)a->nnz;
        p.n = n;
        p.ldb = (uint32_t)b->ld;
        p.ldc = (uint32_t)c->ld;
        p.ldcar = (uint32_t)ldcar;
        p.items = items;
        p.num_chunks = num_chunks;
        p.flags = flags;
        const size_t smem = merge_kernel_smem_bytes(a->dtype, sh, block, items);
        if (smem > (size_t)g_rt.max_smem_optin - 1024) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_merge: items do not fit shared memory");
        int grid = 0;
        BSM_TRY(launch_spmm_merge(a->dtype, sh, p, block, smem, g_rt.stream, &grid));
        int fix_launches = 0;
        BSM_TRY(launch_merge_fixup(a->dtype, p, g_rt.stream, &fix_launches));
        g_info.kernels += 1 + fix_launches;
        g_info.vec_elems = sh.V;
        g_info.lanes_per_row = sh.G;
        g_info.reg_tiles = sh.NT;
        g_info.grid = grid;
        g_info.block = block;
        g_info.smem_bytes = (int)smem;
        g_info.merge_items = (int)items;
        g_info.merge_chunks = (int)num_chunks;
    }
    g_info.passes = passes;
    return BSM_OK;
}

// Row-block probe of a handle (cached): is every row a run of consecutive columns, and how many B rows would the
// row-block kernel load in all?
static int ensure_rowblock_probe(bsm_csr *a)
{
    if (a->rowblock_state) return BSM_OK;
    unsigned long long *d = nullptr, h[2] = {0, 0};
    BSM_CUDA(cudaMallocAsync(&d, 16, g_rt.stream));
    BSM_CUDA(cudaMemsetAsync(d, 0, 16, g_rt.stream));
    int st = launch_rowblock_probe(a->row_ptr, a->col_idx, a->rows, d, g_rt.stream);
    if (st == BSM_OK && cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, g_rt.stream) != cudaSuccess) st = fail(BSM_ERR_CUDA, "rowblock probe: copy failed");
    cudaFreeAsync(d, g_rt.stream);
    BSM_TRY(st);
    BSM_CUDA(cudaStreamSynchronize(g_rt.stream));
    a->rowblock_union = h[0];
    a->rowblock_state = h[1] ? 2 : 1;
    g_info.kernels += 1;
    return BSM_OK;
}

// vector CSR for band-like matrices: blocks of kRowBlockRows consecutive rows share their B-row loads
static int spmm_rowblock(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning &tn, uint32_t flags)
{
    const size_t s = dtype_size(a->dtype);
    const uint32_t n_total = (uint32_t)b->cols;
    const int vmax = (int)(16 / s);
    uint32_t tile = tn.col_tile > 0 ? (uint32_t)tn.col_tile : n_total;
    tile = std::min<uint32_t>(tile, 32u * vmax);   // one register tile per lane
    if (tile < n_total && tile % vmax) tile = std::max<uint32_t>(vmax, tile / vmax * vmax);
    const int probe_kernels = g_info.kernels;
    g_info = bsm_launch_info();
    g_info.kernels = probe_kernels;
    g_info.algo = BSM_ALGO_ROWBLOCK;
    g_info.col_tile = (int)tile;
    int passes = 0;
    for (uint32_t col0 = 0, n = 0; col0 < n_total; col0 += n, ++passes) {
        n = std::min(tile, n_total - col0);
        Shape sh = pick_shape(n, b->ld, c->ld, col0, b->data, c->data, s, false);
        if (sh.NT > 1) {   // alignment forced narrower vectors: a pass of 32 lanes x V columns
            n = 32u * (uint32_t)sh.V;
            sh = pick_shape(n, b->ld, c->ld, col0, b->data, c->data, s, false);
        }
        RowBlockParams p{};
        p.row_ptr = a->row_ptr;
        p.col_idx = a->col_idx;
        p.vals = a->vals;
        p.B = (const char *)b->data + (size_t)col0 * s;
        p.C = (char *)c->data + (size_t)col0 * s;
        p.rows = (uint32_t)a->rows;
        p.n = n;
        p.ldb = (uint32_t)b->ld;
        p.ldc = (uint32_t)c->ld;
        p.flags = flags;
        int grid = 0, block = 0, smem = 0;
        // rows one lane group accumulates side by side: 8 when a warp holds one block (G = 32), 4 when it holds several
        // (more warps fit; band x32 f32 0.328 -> 0.296 ms, x128 f32 1.26 -> 1.15 ms the other way: r1_sweepag_band_*)
        const int rb = tn.rows_per_slice == 4 || tn.rows_per_slice == 8 ? tn.rows_per_slice : (sh.G == 32 ? 8 : 4);
        BSM_TRY(launch_spmm_rowblock(a->dtype, sh, p, rb, a->max_row_nnz, g_rt.sm_count, (size_t)g_rt.max_smem_optin - 1024, g_rt.stream, &grid, &block, &smem));
        g_info.smem_bytes = smem;
        g_info.kernels += grid > 0;
        g_info.vec_elems = sh.V;
        g_info.lanes_per_row = sh.G;
        g_info.reg_tiles = sh.NT;
        g_info.grid = grid;
        g_info.block = block;
        g_info.rows_per_slice = rb;
    }
    g_info.passes = passes;
    return BSM_OK;
}

// Which kernel family runs a product (`requested` = bsm_algo of the caller, AUTO = the heuristics below).
static int choose_algo(const bsm_csr *a_const, uint64_t n_cols, int requested, int *algo)
{
    bsm_csr *a = const_cast<bsm_csr *>(a_const);   // the probe result is cached in the handle
    if (requested == BSM_ALGO_ROWBLOCK) {
        BSM_TRY(ensure_rowblock_probe(a));
        if (a->rowblock_state != 1)
            return fail(BSM_ERR_NOT_SUPPORTED, "BSM_ALGO_ROWBLOCK: the matrix has a row whose stored columns are not a run of consecutive indices");
        *algo = BSM_ALGO_ROWBLOCK;
        return BSM_OK;
    }
    *algo = requested;
    if (requested == BSM_ALGO_VECTOR || requested == BSM_ALGO_MERGE) return BSM_OK;
    // csr_row_stats heuristic: the vector kernel serialises a row on one lane group, so one row far
    // above the mean (power-law hubs, the bench-as-written matrix) needs the nnz-balanced kernel
    const double mean = mean_row_nnz(a);
    *algo = BSM_ALGO_MERGE;
    if ((double)a->max_row_nnz > 64.0 + 8.0 * mean) return BSM_OK;
    // few, long rows (down to one giant row: a checksum vector, the bench-as-written matrix): fewer rows
    // than the vector kernel has warps, so only an entry-balanced split fills the machine
    if (a->rows < (uint64_t)g_rt.sm_count * 96 && a->max_row_nnz > 1024) return BSM_OK;
    *algo = BSM_ALGO_VECTOR;
    // band-like: long regular rows that are runs of consecutive columns, neighbouring rows sharing most of them
    // (at least 2x fewer B-row loads than entries) -> the row-block variant of the vector kernel
    // (output rows of at least 64 bytes: with fewer lanes per row the blocks' value reads are too scattered — SpMV on the
    // band measured 0.50 vs 0.11 ms)
    if (mean >= 16.0 && (double)a->max_row_nnz <= 4.0 * mean + 8.0 && a->rows >= (uint64_t)g_rt.sm_count * 64 &&
        n_cols * dtype_size(a->dtype) >= 64) {
        BSM_TRY(ensure_rowblock_probe(a));
        if (a->rowblock_state == 1 && a->rowblock_union * 2 <= a->nnz) *algo = BSM_ALGO_ROWBLOCK;
    }
    return BSM_OK;
}

static int spmm_dispatch(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning *tuning)
{
    BSM_TRY(ensure_init());
    if (!a || !b || !c) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm: null handle");
    // src/sparse.rs:427-429
    if (a->cols != b->rows) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "spmm: A.cols != B.rows (MatErr::IncorrectDimensions)");
    if (c->rows != a->rows || c->cols != b->cols)
        return fail(BSM_ERR_INCORRECT_DIMENSIONS, "spmm: C must be A.rows x B.cols");
    if (a->dtype != b->dtype || a->dtype != c->dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "spmm: dtype mismatch");
    if (c->data == b->data && c->rows && c->cols) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm: C must not alias B");
    bsm_tuning tn{};
    if (tuning) tn = *tuning;
    uint32_t flags = tn.flags ? (tn.flags & 0x7FFFFFFFu) : BSM_TUNE_DEFAULT_FLAGS;
    g_info = bsm_launch_info();
    if (a->rows == 0 || b->cols == 0) return BSM_OK;
    int algo = BSM_ALGO_VECTOR;
    BSM_TRY(choose_algo(a, b->cols, tn.algo, &algo));
    if (algo == BSM_ALGO_MERGE) return spmm_merge(a, b, c, tn, flags);
    if (algo == BSM_ALGO_ROWBLOCK) {
        const int st = spmm_rowblock(a, b, c, tn, flags);
        // chosen by the heuristic but a warp block's values do not fit shared memory: the plain vector kernel
        if (st != BSM_ERR_NOT_SUPPORTED || tn.algo == BSM_ALGO_ROWBLOCK) return st;
    }
    return spmm_vector(a, b, c, tn, flags);
}

static int dense_to_csr_impl(const bsm_dense *d, bsm_csr **out)
{
    BSM_TRY(ensure_init());
    if (!d || !out) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_to_csr: null argument");
    cudaStream_t sm = g_rt.stream;
    uint32_t *counts = nullptr;
    unsigned long long *total = nullptr;
    bsm_csr *r = nullptr;
    int st = [&]() -> int {
        BSM_TRY(tmp_alloc((void **)&counts, (pad4(d->rows + 1) + 4) * 4));
        BSM_TRY(tmp_alloc((void **)&total, 8));
        BSM_CUDA(cudaMemsetAsync(total, 0, 8, sm));
        BSM_CUDA(cudaMemsetAsync(counts, 0, (pad4(d->rows + 1) + 4) * 4, sm));
        BSM_TRY(launch_count_nonzero(d->dtype, d->data, d->rows, d->cols, d->ld, counts, total, sm));
        BSM_TRY(exclusive_scan_u32(counts, counts, d->rows + 1, sm));   // row_index incl. the finalise() tail
        unsigned long long h = 0;
        BSM_CUDA(cudaMemcpyAsync(&h, total, 8, cudaMemcpyDeviceToHost, sm));
        BSM_CUDA(cudaStreamSynchronize(sm));
        BSM_TRY(alloc_csr(d->dtype, d->rows, d->cols, h, &r));
        BSM_CUDA(cudaMemcpyAsync(r->row_ptr, counts, (d->rows + 1) * 4, cudaMemcpyDeviceToDevice, sm));
        BSM_TRY(launch_scatter_nonzero(d->dtype, d->data, d->rows, d->cols, d->ld, r->row_ptr, r->vals, r->col_idx, sm));
        r->max_row_nnz = d->cols;
        BSM_CUDA(cudaStreamSynchronize(sm));
        return BSM_OK;
    }();
    tmp_free(counts);
    tmp_free(total);
    if (st != BSM_OK) {
        if (r) bsm_csr_free(r);
        return st;
    }
    *out = r;
    return BSM_OK;
}

template <typename T>
static int mul_dense_host(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const T *v, const uint64_t *col_index,
                          const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                          const T *const *rhs_col_ptrs, int algo, uint64_t *out_nnz, T **out_v, uint64_t **out_col_index,
                          uint64_t **out_row_index)
{
    if (!out_nnz || !out_v || !out_col_index || !out_row_index) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense_host: null out");
    BSM_TRY(ensure_init());
    PoolScope pool;
    if (cols != rhs_rows) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "mul_dense: A.cols != rhs.rows (MatErr::IncorrectDimensions)");
    bsm_csr *a = nullptr, *r = nullptr;
    bsm_dense *b = nullptr, *c = nullptr;
    T *hv = nullptr;
    uint64_t *hc = nullptr, *hr = nullptr;
    int st = [&]() -> int {
        BSM_TRY(csr_upload<T>(dtype, rows, cols, nnz, v, col_index, row_index, row_index_len, &a));
        BSM_TRY(dense_upload<T>(dtype, rhs_rows, rhs_cols, rhs_col_ptrs, &b));
        BSM_TRY(dense_alloc(dtype, rows, rhs_cols, &c));
        BSM_TRY(bsm_spmm(a, b, c, algo));
        BSM_TRY(dense_to_csr_impl(c, &r));
        hv = (T *)malloc(std::max<uint64_t>(1, r->nnz) * sizeof(T));
        hc = (uint64_t *)malloc(std::max<uint64_t>(1, r->nnz) * 8);
        hr = (uint64_t *)malloc((r->rows + 1) * 8);
        if (!hv || !hc || !hr) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense_host: out of host memory");
        BSM_TRY((csr_download<T>(r, dtype, hv, hc, hr)));
        *out_nnz = r->nnz;
        return BSM_OK;
    }();
    if (a) bsm_csr_free(a);
    if (b) bsm_dense_free(b);
    if (c) bsm_dense_free(c);
    if (r) bsm_csr_free(r);
    if (st != BSM_OK) {
        free(hv);
        free(hc);
        free(hr);
        return st;
    }
    *out_v = hv;
    *out_col_index = hc;
    *out_row_index = hr;
    return BSM_OK;
}

// Host-to-host dense product, pipelined over groups of columns: column c of C depends on column c of
// B only, so while group g is multiplied its successor's columns travel host->device and its
// predecessor's result columns travel device->host (PCIe is full duplex). Three streams, two sets of
// staging buffers.
template <typename T>
static int mul_dense_host_dense(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const T *v, const uint64_t *col_index,
                                const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                const T *const *rhs_col_ptrs, T *const *out_col_ptrs, int algo)
{
    if (cols != rhs_rows) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "mul_dense: A.cols != rhs.rows (MatErr::IncorrectDimensions)");
    if (rhs_cols && (!rhs_col_ptrs || !out_col_ptrs)) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense_host_dense: null column pointers");
    BSM_TRY(ensure_init());
    PoolScope pool;
    bsm_csr *a = nullptr;
    BSM_TRY(csr_upload<T>(dtype, rows, cols, nnz, v, col_index, row_index, row_index_len, &a));
    if (rows == 0 || rhs_cols == 0) {
        bsm_csr_free(a);
        return BSM_OK;
    }
    // columns per group: about 1 GB of B per group (deep enough a pipeline at large sizes), 4..32
    uint64_t w = 32;
    while (w > 4 && w * rhs_rows * sizeof(T) > ((uint64_t)1 << 30)) w /= 2;
    w = std::min<uint64_t>(rhs_cols, w);
    const uint64_t ngroups = (rhs_cols + w - 1) / w;
    const uint64_t win_lo = a->nnz ? a->col_min : 0, win_rows = a->nnz ? (uint64_t)a->col_max - a->col_min + 1 : 0;
    cudaStream_t user_stream = g_rt.stream, s_in = nullptr, s_mm = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {}, ev_in_free[2] = {}, ev_c[2] = {}, ev_out_free[2] = {};
    T *stage_in[2] = {}, *stage_out[2] = {}, *bg[2] = {}, *cg[2] = {};
    int st = [&]() -> int {
        BSM_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        BSM_CUDA(cudaStreamCreateWithFlags(&s_mm, cudaStreamNonBlocking));
        BSM_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            BSM_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            BSM_CUDA(cudaEventCreateWithFlags(&ev_in_free[i], cudaEventDisableTiming));
            BSM_CUDA(cudaEventCreateWithFlags(&ev_c[i], cudaEventDisableTiming));
            BSM_CUDA(cudaEventCreateWithFlags(&ev_out_free[i], cudaEventDisableTiming));
            BSM_TRY(tmp_alloc((void **)&stage_in[i], w * rhs_rows * sizeof(T)));
            BSM_TRY(tmp_alloc((void **)&bg[i], w * rhs_rows * sizeof(T) + 16));
            BSM_TRY(tmp_alloc((void **)&cg[i], w * rows * sizeof(T) + 16));
            BSM_TRY(tmp_alloc((void **)&stage_out[i], w * rows * sizeof(T)));
        }
        // the pool hands the buffers out in the order of the library stream: make the three pipeline
        // streams wait for that point
        BSM_CUDA(cudaEventRecord(ev_in_free[0], user_stream));
        BSM_CUDA(cudaStreamWaitEvent(s_in, ev_in_free[0], 0));
        BSM_CUDA(cudaStreamWaitEvent(s_mm, ev_in_free[0], 0));
        BSM_CUDA(cudaStreamWaitEvent(s_out, ev_in_free[0], 0));
        for (uint64_t g = 0; g < ngroups; ++g) {
            const int i = (int)(g & 1);
            const uint64_t c0 = g * w, gc = std::min<uint64_t>(w, rhs_cols - c0);
            // host -> device: the group's columns, back to back (column-major staging)
            if (g >= 2) BSM_CUDA(cudaStreamWaitEvent(s_in, ev_in_free[i], 0));
            for (uint64_t c = 0; c < gc; ++c) {
                if (!rhs_col_ptrs[c0 + c] || !out_col_ptrs[c0 + c]) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense_host_dense: null column");
                // only the rows of B that A references travel (a rank's row block of a banded / stencil
                // matrix reads a window of B, not all of it); the rest of the staging is never gathered
                if (win_rows)
                    BSM_CUDA(cudaMemcpyAsync(stage_in[i] + c * rhs_rows + win_lo, rhs_col_ptrs[c0 + c] + win_lo, win_rows * sizeof(T),
                                             cudaMemcpyHostToDevice, s_in));
            }
            BSM_CUDA(cudaEventRecord(ev_in[i], s_in));
            // multiply: column-major -> row-major, C_g = A * B_g, row-major -> column-major
            BSM_CUDA(cudaStreamWaitEvent(s_mm, ev_in[i], 0));
            if (g >= 2) BSM_CUDA(cudaStreamWaitEvent(s_mm, ev_out_free[i], 0));
            BSM_TRY(launch_transpose_cm2rm(dtype, stage_in[i], bg[i], rhs_rows, gc, gc, s_mm));
            BSM_CUDA(cudaEventRecord(ev_in_free[i], s_mm));
            bsm_dense bd, cd;
            bd.dtype = cd.dtype = dtype;
            bd.rows = rhs_rows;
            cd.rows = rows;
            bd.cols = cd.cols = bd.ld = cd.ld = gc;
            bd.data = bg[i];
            cd.data = cg[i];
            bd.owns = cd.owns = false;
            g_rt.stream = s_mm;
            const int rc = bsm_spmm(a, &bd, &cd, algo);
            g_rt.stream = user_stream;
            BSM_TRY(rc);
            BSM_TRY(launch_transpose_rm2cm(dtype, cg[i], stage_out[i], rows, gc, gc, s_mm));
            BSM_CUDA(cudaEventRecord(ev_c[i], s_mm));
            // device -> host: the group's result columns
            BSM_CUDA(cudaStreamWaitEvent(s_out, ev_c[i], 0));
            for (uint64_t c = 0; c < gc; ++c)
                BSM_CUDA(cudaMemcpyAsync(out_col_ptrs[c0 + c], stage_out[i] + c * rows, rows * sizeof(T), cudaMemcpyDeviceToHost, s_out));
            BSM_CUDA(cudaEventRecord(ev_out_free[i], s_out));
        }
        BSM_CUDA(cudaStreamSynchronize(s_out));
        BSM_CUDA(cudaStreamSynchronize(s_mm));
        return BSM_OK;
    }();
    g_rt.stream = user_stream;
    if (st != BSM_OK) cudaDeviceSynchronize();   // nothing may still be using the buffers freed below
    for (int i = 0; i < 2; ++i) {
        tmp_free(stage_in[i]);
        tmp_free(stage_out[i]);
        tmp_free(bg[i]);
        tmp_free(cg[i]);
        if (ev_in[i]) cudaEventDestroy(ev_in[i]);
        if (ev_in_free[i]) cudaEventDestroy(ev_in_free[i]);
        if (ev_c[i]) cudaEventDestroy(ev_c[i]);
        if (ev_out_free[i]) cudaEventDestroy(ev_out_free[i]);
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_mm) cudaStreamDestroy(s_mm);
    if (s_out) cudaStreamDestroy(s_out);
    bsm_csr_free(a);
    return st;
}

template <typename T>
static int mul_vector(const bsm_csr *a, int dtype, const T *rhs, uint64_t rhs_len, T *out, uint64_t out_len)
{
    BSM_TRY(ensure_init());
    if (!a) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_vector: null handle");
    if (a->dtype != dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "mul_vector: dtype mismatch");
    // src/sparse.rs:469-471
    if (a->cols != rhs_len || a->rows != out_len)
        return fail(BSM_ERR_INCORRECT_DIMENSIONS, "mul_vector: dims (MatErr::IncorrectDimensions)");
    bsm_dense *x = nullptr, *y = nullptr;
    PoolScope pool;
    int st = [&]() -> int {
        BSM_TRY(dense_alloc(dtype, rhs_len, 1, &x));
        BSM_TRY(dense_alloc(dtype, out_len, 1, &y));
        if (rhs_len) BSM_CUDA(cudaMemcpyAsync(x->data, rhs, rhs_len * sizeof(T), cudaMemcpyHostToDevice, g_rt.stream));
        BSM_TRY(bsm_spmm(a, x, y, BSM_ALGO_AUTO));
        if (out_len) BSM_CUDA(cudaMemcpyAsync(out, y->data, out_len * sizeof(T), cudaMemcpyDeviceToHost, g_rt.stream));
        BSM_CUDA(cudaStreamSynchronize(g_rt.stream));
        return BSM_OK;
    }();
    if (x) bsm_dense_free(x);
    if (y) bsm_dense_free(y);
    return st;
}

template <typename CountFn, typename FillFn>
static int gen_counted(int dtype, uint64_t rows, uint64_t cols, uint64_t max_per_row, uint64_t row_begin, CountFn count_fn, FillFn fill_fn, bsm_csr **out)
{
    BSM_TRY(ensure_init());
    if (!out) return fail(BSM_ERR_INVALID_ARGUMENT, "gen: null out");
    if (dtype != BSM_F32 && dtype != BSM_F64) return fail(BSM_ERR_DTYPE_MISMATCH, "gen: dtype must be f32 or f64");
    if (rows >= 0xFFFFFFF0ull) return fail(BSM_ERR_INDEX_OVERFLOW, "gen: too many rows");
    // the row counts are scanned in u32: refuse anything whose entry count could wrap
    if ((__uint128_t)rows * max_per_row >= 0xFFFFFFF0ull) return fail(BSM_ERR_INDEX_OVERFLOW, "gen: nnz would not fit the device's u32 indices");
    cudaStream_t sm = g_rt.stream;
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
        return compute_stats(a);
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

// ==========================================================================================
// extern "C"
// ==========================================================================================
extern "C" {

int bsm_abi_version(void) { return BSM_ABI_VERSION; }

int bsm_device_count(int *count)
{
    if (!count) return fail(BSM_ERR_INVALID_ARGUMENT, "device_count: null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(BSM_ERR_NO_DEVICE, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    }
    *count = n;
    return BSM_OK;
}

int bsm_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(BSM_ERR_NO_DEVICE, std::string("no usable CUDA device (this library has no CPU fallback): ") +
                                           (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(BSM_ERR_INVALID_ARGUMENT, "bsm_init: device index out of range");
    BSM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    BSM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (g_rt.own_stream && g_rt.device != device) {
        cudaStreamDestroy(g_rt.own_stream);
        g_rt.own_stream = nullptr;
    }
    if (!g_rt.own_stream) BSM_CUDA(cudaStreamCreateWithFlags(&g_rt.own_stream, cudaStreamNonBlocking));
    g_rt.stream = g_rt.own_stream;
    g_rt.device = device;
    g_rt.sm_count = prop.multiProcessorCount;
    g_rt.l2_bytes = (size_t)prop.l2CacheSize;
    g_rt.hbm_bytes = prop.totalGlobalMem;
    g_rt.cc_major = prop.major;
    g_rt.cc_minor = prop.minor;
    g_rt.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
        unsigned long long keep = ~0ull;   // never trim the stream-ordered pool behind our back
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return BSM_OK;
}

int bsm_set_stream(void *cuda_stream)
{
    BSM_TRY(ensure_init());
    g_rt.stream = cuda_stream ? (cudaStream_t)cuda_stream : g_rt.own_stream;
    return BSM_OK;
}

int bsm_sync(void)
{
    BSM_TRY(ensure_init());
    BSM_CUDA(cudaStreamSynchronize(g_rt.stream));
    return BSM_OK;
}

const char *bsm_last_error_string(void) { return g_err.c_str(); }

const char *bsm_status_string(int status)
{
    switch (status) {
        case BSM_OK: return "ok";
        case BSM_ERR_INCORRECT_DIMENSIONS: return "IncorrectDimensions";
        case BSM_ERR_NOT_FINALISED: return "MatrixNotFinalised";
        case BSM_ERR_OUT_OF_BOUNDS: return "OutOfBounds";
        case BSM_ERR_INDEX_OVERFLOW: return "IndexOverflow";
        case BSM_ERR_INVALID_ARGUMENT: return "InvalidArgument";
        case BSM_ERR_DTYPE_MISMATCH: return "DtypeMismatch";
        case BSM_ERR_CUDA: return "CudaError";
        case BSM_ERR_NCCL: return "NcclError";
        case BSM_ERR_NO_DEVICE: return "NoDevice";
        case BSM_ERR_NOT_SUPPORTED: return "NotSupported";
    }
    return "unknown";
}

int bsm_device_info(int *sm_count, size_t *l2_bytes, size_t *hbm_bytes, int *cc_major, int *cc_minor)
{
    BSM_TRY(ensure_init());
    if (sm_count) *sm_count = g_rt.sm_count;
    if (l2_bytes) *l2_bytes = g_rt.l2_bytes;
    if (hbm_bytes) *hbm_bytes = g_rt.hbm_bytes;
    if (cc_major) *cc_major = g_rt.cc_major;
    if (cc_minor) *cc_minor = g_rt.cc_minor;
    return BSM_OK;
}

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
        cudaStream_t sm = g_rt.stream;
        cudaMemcpyAsync(a->vals, d_vals, nnz * dtype_size(dtype), cudaMemcpyDeviceToDevice, sm);
        cudaMemcpyAsync(a->col_idx, d_col_idx, nnz * 4, cudaMemcpyDeviceToDevice, sm);
        cudaMemcpyAsync(a->row_ptr, d_row_ptr, (rows + 1) * 4, cudaMemcpyDeviceToDevice, sm);
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
    int st = compute_stats(a);
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
    if (bytes) BSM_CUDA(cudaMemsetAsync(d->data, 0, bytes, g_rt.stream));
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
    BSM_CUDA(cudaMemcpy2DAsync(dst, d->cols * s, d->data, d->ld * s, d->cols * s, d->rows, cudaMemcpyDeviceToHost, g_rt.stream));
    BSM_CUDA(cudaStreamSynchronize(g_rt.stream));
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
        cudaError_t e = cudaMemcpy2DAsync(d->data, d->ld * s, src, cols * s, cols * s, rows, cudaMemcpyHostToDevice, g_rt.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_rt.stream);
        if (e != cudaSuccess) {
            bsm_dense_free(d);
            return fail(BSM_ERR_CUDA, std::string("dense_upload_rowmajor: ") + cudaGetErrorString(e));
        }
    }
    *out = d;
    return BSM_OK;
}

// ---- hot path ------------------------------------------------------------------------------------
int bsm_spmm(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, int algo)
{
    bsm_tuning tn{};
    tn.algo = algo;
    return spmm_dispatch(a, b, c, &tn);
}
int bsm_spmm_tuned(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning *tuning)
{
    return spmm_dispatch(a, b, c, tuning);
}
int bsm_spmm_scatter(const bsm_csr *a, const bsm_dense *b, bsm_dense *const *c_full, int ndest, uint64_t row_offset, int algo)
{
    BSM_TRY(ensure_init());
    if (!a || !b || !c_full || ndest < 1 || ndest > 8) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_scatter: bad arguments (1..8 destinations)");
    if (a->cols != b->rows) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "spmm_scatter: A.cols != B.rows (MatErr::IncorrectDimensions)");
    ScatterTargets st;
    st.row_offset = row_offset;
    for (int d = 0; d < ndest; ++d) {
        const bsm_dense *c = c_full[d];
        if (!c) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_scatter: null destination");
        if (c->dtype != a->dtype || b->dtype != a->dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "spmm_scatter: dtype mismatch");
        if (c->cols != b->cols || c->rows < row_offset + a->rows || c->ld != c_full[0]->ld)
            return fail(BSM_ERR_INCORRECT_DIMENSIONS, "spmm_scatter: every destination must be (>= row_offset + A.rows) x B.cols with one leading dimension");
        if (c->data == b->data) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_scatter: a destination aliases B");
        if (d > 0) st.peer_data[st.n_peers++] = c->data;
    }
    // a row may be summed by one lane group only (its row is written, never read back): vector kernel
    int chosen = algo;
    if (algo == BSM_ALGO_AUTO) BSM_TRY(choose_algo(a, b->cols, algo, &chosen));
    if (chosen == BSM_ALGO_MERGE)
        return fail(BSM_ERR_NOT_SUPPORTED, "spmm_scatter: the merge-path kernel revisits C rows (fix-up) and cannot scatter; "
                                           "use bsm_spmm + bsm_allgather_rows for power-law matrices");
    g_info = bsm_launch_info();
    if (a->rows == 0 || b->cols == 0) return BSM_OK;
    bsm_tuning tn{};
    return spmm_vector(a, b, c_full[0], tn, BSM_TUNE_DEFAULT_FLAGS, &st);
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

int bsm_last_launch_info(bsm_launch_info *info)
{
    if (!info) return fail(BSM_ERR_INVALID_ARGUMENT, "last_launch_info: null");
    *info = g_info;
    return BSM_OK;
}
uint64_t bsm_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

uint32_t bsm_line_length_of_row(const uint32_t *cols, uint32_t len, uint64_t diag) { return cols && len ? line_length_of_row(cols, len, diag) : 0u; }

// Dry run of the vector kernel's launch heuristics — pure host arithmetic, no device needed (the CPU test-suite pins
// the heuristics with it). Operands are assumed 16-byte aligned with ld = n rounded up to 16 bytes; `grid` assumes the
// occupancy the chosen flavour targets.
int bsm_plan_vector(int dtype, uint64_t rows, uint64_t nnz, uint64_t max_row_nnz, uint32_t row_stride, uint64_t n_cols,
                    const bsm_tuning *tuning, int sm_count, uint64_t smem_optin_bytes, bsm_launch_info *out)
{
    if (!out) return fail(BSM_ERR_INVALID_ARGUMENT, "plan_vector: null out");
    if (dtype != BSM_F32 && dtype != BSM_F64) return fail(BSM_ERR_DTYPE_MISMATCH, "plan_vector: dtype must be f32 or f64");
    if (sm_count <= 0 || smem_optin_bytes < 2048 || n_cols == 0 || n_cols >= 0xFFFFFFF0ull)
        return fail(BSM_ERR_INVALID_ARGUMENT, "plan_vector: bad device facts or column count");
    bsm_tuning tn{};
    if (tuning) tn = *tuning;
    const size_t s = dtype_size(dtype);
    const int vmax = (int)(16 / s);
    const uint32_t n_total = (uint32_t)n_cols;
    uint32_t tile = tn.col_tile > 0 ? (uint32_t)tn.col_tile : n_total;
    tile = std::min<uint32_t>(tile, 32u * vmax * 4u);
    if (tile < n_total && tile % vmax) tile = std::max<uint32_t>(vmax, tile / vmax * vmax);
    const MatrixFacts m{dtype, rows, nnz, max_row_nnz, row_stride};
    const DeviceFacts dev{sm_count, (size_t)smem_optin_bytes - 1024};
    const uint64_t ld = round_up(n_total, vmax);
    // the plan reported is that of the FIRST pass; `passes` counts them all, `capacity`... see bsm_launch_info
    uint32_t first_n = 0;
    int passes = 0;
    for (uint32_t col0 = 0, n = 0; col0 < n_total; col0 += n, ++passes) {
        n = fit_pass_width(std::min(tile, n_total - col0), vmax, [&](uint32_t w) { return pick_shape(w, ld, ld, col0, nullptr, nullptr, s, tn.prefer_wide_rows != 0); });
        if (n == 0) return fail(BSM_ERR_INVALID_ARGUMENT, "plan_vector: empty pass");
        if (!first_n) first_n = n;
    }
    VectorPlan plan;
    BSM_TRY(plan_vector_pass(m, dev, tn, first_n, PassAlign{ld, ld, 0, nullptr, nullptr}, false, false, &plan));
    *out = bsm_launch_info();
    out->algo = BSM_ALGO_VECTOR;
    out->vec_elems = plan.sh.V;
    out->lanes_per_row = plan.sh.G;
    out->reg_tiles = plan.sh.NT;
    out->block = plan.nw * 32;
    out->grid = (int)std::min<uint64_t>(plan.num_super, (uint64_t)sm_count * plan.resident);
    out->smem_bytes = (int)plan.smem;
    out->rows_per_slice = (int)plan.R;
    out->rows_per_warp = (int)plan.P;
    out->stages = (int)plan.stages;
    out->capacity = (int)plan.cap;
    out->reg_flavour = plan.flavour + 1;
    out->col_tile = (int)first_n;   // width of the first pass
    out->passes = passes;
    return BSM_OK;
}

int bsm_dense_residual_norm(const bsm_dense *ax, const bsm_dense *b, double *resid_fro, double *b_fro)
{
    BSM_TRY(ensure_init());
    if (!ax || !b || !resid_fro || !b_fro) return fail(BSM_ERR_INVALID_ARGUMENT, "residual_norm: null argument");
    if (ax->rows != b->rows || ax->cols != b->cols) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "residual_norm: shapes differ");
    if (ax->dtype != b->dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "residual_norm: dtype mismatch");
    const int nd = residual_norm_scratch_doubles();
    double *d = nullptr;
    BSM_TRY(tmp_alloc((void **)&d, nd * sizeof(double)));
    int st = launch_residual_norms(ax->dtype, ax->data, ax->ld, b->data, b->ld, ax->rows, ax->cols, d, g_rt.stream);
    double h[2] = {0.0, 0.0};
    if (st == BSM_OK) {
        cudaError_t e = cudaMemcpyAsync(h, d + nd - 2, 16, cudaMemcpyDeviceToHost, g_rt.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_rt.stream);
        if (e != cudaSuccess) st = fail(BSM_ERR_CUDA, std::string("residual_norm: ") + cudaGetErrorString(e));
    }
    tmp_free(d);
    if (st != BSM_OK) return st;
    *resid_fro = std::sqrt(h[0]);
    *b_fro = std::sqrt(h[1]);
    return BSM_OK;
}

int bsm_dense_to_csr(const bsm_dense *d, bsm_csr **out) { return dense_to_csr_impl(d, out); }

int bsm_mul_dense_host_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v, const uint64_t *col_index,
                           const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                           const double *const *rhs_col_ptrs, int algo, uint64_t *out_nnz, double **out_v,
                           uint64_t **out_col_index, uint64_t **out_row_index)
{
    return mul_dense_host<double>(BSM_F64, rows, cols, nnz, v, col_index, row_index, row_index_len, rhs_rows, rhs_cols,
                                  rhs_col_ptrs, algo, out_nnz, out_v, out_col_index, out_row_index);
}
int bsm_mul_dense_host_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v, const uint64_t *col_index,
                           const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                           const float *const *rhs_col_ptrs, int algo, uint64_t *out_nnz, float **out_v,
                           uint64_t **out_col_index, uint64_t **out_row_index)
{
    return mul_dense_host<float>(BSM_F32, rows, cols, nnz, v, col_index, row_index, row_index_len, rhs_rows, rhs_cols,
                                 rhs_col_ptrs, algo, out_nnz, out_v, out_col_index, out_row_index);
}
int bsm_mul_dense_host_dense_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v, const uint64_t *col_index,
                                 const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                 const double *const *rhs_col_ptrs, double *const *out_col_ptrs, int algo)
{
    return mul_dense_host_dense<double>(BSM_F64, rows, cols, nnz, v, col_index, row_index, row_index_len, rhs_rows, rhs_cols,
                                        rhs_col_ptrs, out_col_ptrs, algo);
}
int bsm_mul_dense_host_dense_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v, const uint64_t *col_index,
                                 const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                 const float *const *rhs_col_ptrs, float *const *out_col_ptrs, int algo)
{
    return mul_dense_host_dense<float>(BSM_F32, rows, cols, nnz, v, col_index, row_index, row_index_len, rhs_rows, rhs_cols,
                                       rhs_col_ptrs, out_col_ptrs, algo);
}
void bsm_host_free(void *p) { free(p); }

int bsm_mul_vector_f64(const bsm_csr *a, const double *rhs, uint64_t rhs_len, double *out, uint64_t out_len)
{
    return mul_vector<double>(a, BSM_F64, rhs, rhs_len, out, out_len);
}
int bsm_mul_vector_f32(const bsm_csr *a, const float *rhs, uint64_t rhs_len, float *out, uint64_t out_len)
{
    return mul_vector<float>(a, BSM_F32, rhs, rhs_len, out, out_len);
}

// ---- sharding ------------------------------------------------------------------------------------
int bsm_partition_rows(const uint64_t *row_index, uint64_t rows, int parts, uint64_t *bounds)
{
    if (!row_index || !bounds || parts < 1) return fail(BSM_ERR_INVALID_ARGUMENT, "partition_rows: bad arguments");
    const uint64_t nnz = row_index[rows] - row_index[0];
    bounds[0] = 0;
    for (int p = 1; p < parts; ++p) {
        // first row whose start offset reaches p/parts of the entries (ties -> equal row counts)
        const uint64_t target = row_index[0] + (uint64_t)((__uint128_t)nnz * (unsigned)p / (unsigned)parts);
        uint64_t r;
        if (nnz == 0) {
            r = rows * (uint64_t)p / (uint64_t)parts;
        } else {
            r = (uint64_t)(std::lower_bound(row_index, row_index + rows + 1, target) - row_index);
            if (r > rows) r = rows;
        }
        bounds[p] = std::max(r, bounds[p - 1]);
    }
    bounds[parts] = rows;
    return BSM_OK;
}

// ---- generators ------------------------------------------------------------------------------------
int bsm_gen_dense(int dtype, uint64_t rows, uint64_t cols, uint64_t seed, int mode, double offset, bsm_dense **out)
{
    bsm_dense *d = nullptr;
    BSM_TRY(dense_alloc(dtype, rows, cols, &d));
    int st = launch_gen_dense(dtype, d->data, rows, cols, d->ld, seed, mode, offset, g_rt.stream);
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
    int st = gen_rmat_device(dtype, scale, edges, pa, pb, pc, seed, mode, a->row_ptr, a->col_idx, a->vals, g_rt.stream);
    if (st == BSM_OK) st = compute_stats(a);
    if (st != BSM_OK) {
        bsm_csr_free(a);
        return st;
    }
    *out = a;
    return BSM_OK;
}

int bsm_l2_flush(void)
{
    BSM_TRY(ensure_init());
    const size_t want = std::max<size_t>(g_rt.l2_bytes * 2, (size_t)256 << 20);
    if (g_rt.flush_bytes < want) {
        if (g_rt.flush_buf) cudaFree(g_rt.flush_buf);
        g_rt.flush_buf = nullptr;
        g_rt.flush_bytes = 0;
        BSM_CUDA(cudaMalloc(&g_rt.flush_buf, want));
        g_rt.flush_bytes = want;
    }
    BSM_CUDA(cudaMemsetAsync(g_rt.flush_buf, 0, g_rt.flush_bytes, g_rt.stream));
    return BSM_OK;
}

}  // extern "C"
