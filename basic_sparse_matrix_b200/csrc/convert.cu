// convert.cu — format conversion on either side of the hot path (sm_100a):
//   * usize (u64) -> u32 index narrowing with bound check   (Csr fields, src/sparse.rs:72-73)
//   * column-major Vec<Vec<T>> <-> row-major device Dense    (src/dense.rs:5-9)
//   * csr_row_stats (max row length -> kernel dispatch)
//   * dense -> Csr zero-drop compaction: count -> scan -> scatter
//     (result construction of mul_dense: sparse.rs:442 -> insert 222-233 -> finalise 206-219)
#include <algorithm>

#include "bsm_common.cuh"
#include "kernels.h"
#include "line_length.h"

namespace bsm {

// ---- index narrowing / widening --------------------------------------------------------------
__global__ void narrow_u64_kernel(const uint64_t *__restrict__ src, uint32_t *__restrict__ dst, uint64_t count,
                                  uint64_t bound, uint64_t subtract, uint32_t *flag)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint64_t v = src[i] - subtract;
        bad |= v >= bound;
        dst[i] = (uint32_t)v;
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 1u);
}

__global__ void widen_u32_kernel(const uint32_t *__restrict__ src, uint64_t *__restrict__ dst, uint64_t count)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = src[i];
}

__global__ void fill_u32_kernel(uint32_t *dst, uint64_t count, uint32_t value)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = value;
}

// grid-stride launches: enough blocks to fill the device the runtime was initialised on (16 per SM)
static inline int grid_for(uint64_t count, int threads)
{
    const int sms = rt().sm_count > 0 ? rt().sm_count : 148;
    const uint64_t max_blocks = (uint64_t)sms * 16;
    uint64_t b = (count + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

int launch_narrow_u64(const uint64_t *src, uint32_t *dst, uint64_t count, uint64_t bound, uint64_t subtract,
                      uint32_t *flag, cudaStream_t stream)
{
    if (count == 0) return BSM_OK;
    narrow_u64_kernel<<<grid_for(count, 256), 256, 0, stream>>>(src, dst, count, bound, subtract, flag);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

int launch_widen_u32(const uint32_t *src, uint64_t *dst, uint64_t count, cudaStream_t stream)
{
    if (count == 0) return BSM_OK;
    widen_u32_kernel<<<grid_for(count, 256), 256, 0, stream>>>(src, dst, count);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

int launch_fill_u32(uint32_t *dst, uint64_t count, uint32_t value, cudaStream_t stream)
{
    if (count == 0) return BSM_OK;
    fill_u32_kernel<<<grid_for(count, 256), 256, 0, stream>>>(dst, count, value);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

// ---- per-matrix statistics: ONE kernel, one readback ---------------------------------------------------------
// stats[kStatMaxLen] longest row, [kStatBadRowPtr] row_ptr not non-decreasing / not ending at nnz / not starting at 0,
// [kStatColMin/Max] smallest and largest stored column (which rows of B the matrix — or a rank's row block — reads),
// [kStatVoters] rows of 3..64 entries, [kStatHist + s] how many of them suggest the stencil line length s
// (line_length.h): the host takes the dominant one. Every row votes (warp-aggregated atomics), not a sample of three.
__global__ void csr_stats_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col_idx, uint64_t rows, uint64_t nnz,
                                 uint64_t row_offset, uint32_t *__restrict__ stats)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    // entries: column range
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (uint64_t i = tid; i < nnz; i += stride) {
        const uint32_t c = col_idx[i];
        lo = min(lo, c);
        hi = max(hi, c);
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
    }
    if (lane == 0 && lo <= hi) {
        atomicMin(stats + kStatColMin, lo);
        atomicMax(stats + kStatColMax, hi);
    }
    // rows: longest row, monotonicity, line-length votes
    uint32_t m = 0;
    bool bad = false;
    if (tid == 0 && rows) bad = row_ptr[0] != 0u || row_ptr[rows] != (uint32_t)nnz;
    const uint64_t rows_pad = (rows + 31) / 32 * 32;   // whole warps stay in the loop (warp-wide votes below)
    for (uint64_t i = tid; i < rows_pad; i += stride) {
        uint32_t vote = 0;
        if (i < rows) {
            const uint32_t a = row_ptr[i], b = row_ptr[i + 1];
            if (b < a || b > nnz) {
                bad = true;
            } else {
                const uint32_t len = b - a;
                m = max(m, len);
                if (len >= 3 && len <= 64) {
                    const uint32_t s = line_length_of_row(col_idx + a, len, i + row_offset);
                    vote = (s >= kStrideMin && s <= kStrideMax) ? s : 0xFFFFFFFFu;   // all-ones: voted "no line"
                }
            }
        }
        // warp-aggregated histogram update: one atomic per distinct vote in the warp
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, vote);
        if (vote && lane == (uint32_t)(__ffs(peers) - 1)) {
            atomicAdd(stats + kStatVoters, (uint32_t)__popc(peers));
            if (vote != 0xFFFFFFFFu) atomicAdd(stats + kStatHist + vote, (uint32_t)__popc(peers));
        }
    }
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if (lane == 0 && m) atomicMax(stats + kStatMaxLen, m);
    if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) atomicExch(stats + kStatBadRowPtr, 1u);
}

int launch_csr_stats(const uint32_t *row_ptr, const uint32_t *col_idx, uint64_t rows, uint64_t nnz, uint64_t row_offset, uint32_t *stats,
                     cudaStream_t stream)
{
    csr_stats_kernel<<<grid_for(std::max(rows, nnz), 256), 256, 0, stream>>>(row_ptr, col_idx, rows, nnz, row_offset, stats);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

// ---- residual norms: ||X - Y||_F^2 and ||Y||_F^2 (f64 accumulation, fixed grid -> deterministic) ----
constexpr int kNormBlocks = 1184, kNormThreads = 256;

template <typename T>
__global__ void residual_partial_kernel(const T *__restrict__ x, uint64_t ldx, const T *__restrict__ y, uint64_t ldy, uint64_t rows,
                                        uint64_t cols, double *__restrict__ partial)
{
    __shared__ double sh[2][kNormThreads / 32];
    double d2 = 0.0, y2 = 0.0;
    const uint64_t total = rows * cols;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / cols, c = i - r * cols;
        const double xv = (double)x[r * ldx + c], yv = (double)y[r * ldy + c];
        d2 += (xv - yv) * (xv - yv);
        y2 += yv * yv;
    }
    for (int o = 16; o; o >>= 1) {
        d2 += __shfl_xor_sync(0xFFFFFFFFu, d2, o);
        y2 += __shfl_xor_sync(0xFFFFFFFFu, y2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = d2;
        sh[1][threadIdx.x >> 5] = y2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < kNormThreads / 32; ++w) {
            a += sh[0][w];
            b += sh[1][w];
        }
        partial[2 * blockIdx.x] = a;
        partial[2 * blockIdx.x + 1] = b;
    }
}

__global__ void residual_final_kernel(const double *__restrict__ partial, int blocks, double *__restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < blocks; ++i) {
            a += partial[2 * i];
            b += partial[2 * i + 1];
        }
        out[0] = a;
        out[1] = b;
    }
}

int launch_residual_norms(int dtype, const void *x, uint64_t ldx, const void *y, uint64_t ldy, uint64_t rows, uint64_t cols,
                          double *partial /* [2*kNormBlocks + 2] */, cudaStream_t stream)
{
    if (dtype == BSM_F64)
        residual_partial_kernel<double><<<kNormBlocks, kNormThreads, 0, stream>>>((const double *)x, ldx, (const double *)y, ldy, rows, cols, partial);
    else
        residual_partial_kernel<float><<<kNormBlocks, kNormThreads, 0, stream>>>((const float *)x, ldx, (const float *)y, ldy, rows, cols, partial);
    BSM_CUDA(cudaGetLastError());
    residual_final_kernel<<<1, 32, 0, stream>>>(partial, kNormBlocks, partial + 2 * kNormBlocks);
    BSM_CUDA(cudaGetLastError());
    count_launch(2);
    return BSM_OK;
}
int residual_norm_scratch_doubles() { return 2 * kNormBlocks + 2; }

// ---- layout transposes ---------------------------------------------------------------------------
// colmajor[c*rows + r]  <->  rowmajor[r*ld + c]; 32x32 tiles through padded shared memory so both
// sides are coalesced.
template <typename T, bool TO_ROWMAJOR>
__global__ void transpose_kernel(const T *__restrict__ src, T *__restrict__ dst, uint64_t rows, uint64_t cols, uint64_t ld)
{
    __shared__ T tile[32][33];
    const uint64_t tiles_c = (cols + 31) / 32;
    const uint64_t tiles_r = (rows + 31) / 32;
    const uint64_t ntiles = tiles_c * tiles_r;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
        if (TO_ROWMAJOR) {
            // read column-major: threadIdx.x walks rows (contiguous in a column)
            for (int j = threadIdx.y; j < 32; j += blockDim.y) {
                const uint64_t c = c0 + j, r = r0 + threadIdx.x;
                if (c < cols && r < rows) tile[j][threadIdx.x] = src[c * rows + r];
            }
            __syncthreads();
            for (int j = threadIdx.y; j < 32; j += blockDim.y) {
                const uint64_t r = r0 + j, c = c0 + threadIdx.x;
                if (c < cols && r < rows) dst[r * ld + c] = tile[threadIdx.x][j];
            }
        } else {
            for (int j = threadIdx.y; j < 32; j += blockDim.y) {
                const uint64_t r = r0 + j, c = c0 + threadIdx.x;
                if (c < cols && r < rows) tile[j][threadIdx.x] = src[r * ld + c];
            }
            __syncthreads();
            for (int j = threadIdx.y; j < 32; j += blockDim.y) {
                const uint64_t c = c0 + j, r = r0 + threadIdx.x;
                if (c < cols && r < rows) dst[c * rows + r] = tile[threadIdx.x][j];
            }
        }
        __syncthreads();
    }
}

template <bool TO_ROWMAJOR>
static int launch_transpose(int dtype, const void *src, void *dst, uint64_t rows, uint64_t cols, uint64_t ld,
                            cudaStream_t stream)
{
    if (rows == 0 || cols == 0) return BSM_OK;
    const uint64_t ntiles = ((cols + 31) / 32) * ((rows + 31) / 32);
    const uint64_t max_grid = (uint64_t)(rt().sm_count > 0 ? rt().sm_count : 148) * 32;
    const int grid = (int)(ntiles < max_grid ? ntiles : max_grid);
    dim3 block(32, 8);
    if (dtype == BSM_F64)
        transpose_kernel<double, TO_ROWMAJOR><<<grid, block, 0, stream>>>((const double *)src, (double *)dst, rows, cols, ld);
    else
        transpose_kernel<float, TO_ROWMAJOR><<<grid, block, 0, stream>>>((const float *)src, (float *)dst, rows, cols, ld);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

int launch_transpose_cm2rm(int dtype, const void *colmajor, void *rowmajor, uint64_t rows, uint64_t cols, uint64_t ld,
                           cudaStream_t stream)
{
    return launch_transpose<true>(dtype, colmajor, rowmajor, rows, cols, ld, stream);
}
int launch_transpose_rm2cm(int dtype, const void *rowmajor, void *colmajor, uint64_t rows, uint64_t cols, uint64_t ld,
                           cudaStream_t stream)
{
    return launch_transpose<false>(dtype, rowmajor, colmajor, rows, cols, ld, stream);
}

// ---- exclusive scan (u32), three-phase, recursive over block sums -------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 8;
constexpr int kScanTile = kScanThreads * kScanPerThread;   // 2048

__global__ void scan_tile_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t count,
                                 uint32_t *__restrict__ tile_sums)
{
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPerThread;
    uint32_t v[kScanPerThread];
    uint32_t local = 0;
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i) {
        v[i] = base + i < count ? in[base + i] : 0u;
        local += v[i];
    }
    // exclusive scan of `local` across the block
    uint32_t incl = local;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += n;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0u;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, w, o);
            if (lane >= (uint32_t)o) w += n;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    uint32_t run = incl - local + (warp ? warp_sums[warp - 1] : 0u);
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i) {
        if (base + i < count) out[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == kScanThreads - 1 && tile_sums) tile_sums[blockIdx.x] = run;
}

__global__ void scan_add_kernel(uint32_t *__restrict__ out, uint64_t count, const uint32_t *__restrict__ tile_offsets)
{
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPerThread;
    const uint32_t off = tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i)
        if (base + i < count) out[base + i] += off;
}

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t count, cudaStream_t stream)
{
    if (count == 0) return BSM_OK;
    const uint64_t tiles = (count + kScanTile - 1) / kScanTile;
    uint32_t *tile_sums = nullptr;
    if (tiles > 1) BSM_CUDA(cudaMallocAsync(&tile_sums, tiles * sizeof(uint32_t), stream));
    scan_tile_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, out, count, tile_sums);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    if (tiles > 1) {
        BSM_TRY(exclusive_scan_u32(tile_sums, tile_sums, tiles, stream));
        scan_add_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(out, count, tile_sums);
        BSM_CUDA(cudaGetLastError());
        count_launch();
        BSM_CUDA(cudaFreeAsync(tile_sums, stream));
    }
    return BSM_OK;
}

// ---- dense -> Csr zero-drop compaction ------------------------------------------------------------
// keep(x) == (x != T::default()): NaN kept, -0.0 dropped (src/sparse.rs:229)
template <typename T> __device__ __forceinline__ bool keep(T x) { return x != T(0); }

// one warp per row
// masks (optional): bit c % 64 of masks[r * words + c / 64] = entry (r, c) is kept — the result's column indices in 1 bit instead
// of 8 bytes each (what the pipelined host call sends over PCIe instead of the usize columns)
template <typename T>
__global__ void count_nonzero_kernel(const T *__restrict__ d, uint64_t rows, uint64_t cols, uint64_t ld,
                                     uint32_t *__restrict__ counts, unsigned long long *total, uint64_t *__restrict__ masks)
{
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t words = (cols + 63) / 64;
    unsigned long long local = 0;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        uint32_t n = 0;
        uint64_t w = 0;
        for (uint64_t c0 = 0; c0 < cols; c0 += 32) {
            const uint64_t c = c0 + lane;
            const bool k = c < cols && keep(d[r * ld + c]);
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, k);
            n += __popc(m);
            if (masks) {
                w |= (uint64_t)m << (c0 & 32u);
                if (lane == 0 && ((c0 & 32u) || c0 + 32 >= cols)) {
                    masks[r * words + c0 / 64] = w;
                    w = 0;
                }
            }
        }
        if (lane == 0) {
            counts[r] = n;
            local += n;
        }
    }
    if (total && lane == 0 && local) atomicAdd(total, local);
}

template <typename T>
__global__ void scatter_nonzero_kernel(const T *__restrict__ d, uint64_t rows, uint64_t cols, uint64_t ld,
                                       const uint32_t *__restrict__ row_ptr, T *__restrict__ vals,
                                       uint32_t *__restrict__ col_idx)
{
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        uint32_t pos = row_ptr[r];
        for (uint64_t c0 = 0; c0 < cols; c0 += 32) {
            const uint64_t c = c0 + lane;
            T x = T(0);
            if (c < cols) x = d[r * ld + c];
            const bool k = c < cols && keep(x);
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, k);
            if (k) {
                const uint32_t o = pos + __popc(m & ((1u << lane) - 1u));   // ascending column = insertion order
                vals[o] = x;
                col_idx[o] = (uint32_t)c;
            }
            pos += __popc(m);
        }
    }
}

int launch_count_nonzero(int dtype, const void *dense, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t *counts,
                         unsigned long long *total, cudaStream_t stream, uint64_t *masks)
{
    if (rows == 0) return BSM_OK;
    const int grid = grid_for(rows * 32, 256);
    if (dtype == BSM_F64)
        count_nonzero_kernel<double><<<grid, 256, 0, stream>>>((const double *)dense, rows, cols, ld, counts, total, masks);
    else
        count_nonzero_kernel<float><<<grid, 256, 0, stream>>>((const float *)dense, rows, cols, ld, counts, total, masks);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

int launch_scatter_nonzero(int dtype, const void *dense, uint64_t rows, uint64_t cols, uint64_t ld, const uint32_t *row_ptr,
                           void *vals, uint32_t *col_idx, cudaStream_t stream)
{
    if (rows == 0) return BSM_OK;
    const int grid = grid_for(rows * 32, 256);
    if (dtype == BSM_F64)
        scatter_nonzero_kernel<double><<<grid, 256, 0, stream>>>((const double *)dense, rows, cols, ld, row_ptr, (double *)vals, col_idx);
    else
        scatter_nonzero_kernel<float><<<grid, 256, 0, stream>>>((const float *)dense, rows, cols, ld, row_ptr, (float *)vals, col_idx);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

// ---- pieces of the pipelined host call (pipeline.cu) ------------------------------------------------------------
// smallest / largest stored column of every block of `block_rows` consecutive rows: which window of B a row block reads
__global__ void block_col_range_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col_idx, uint64_t rows,
                                       uint64_t block_rows, uint32_t *__restrict__ out)
{
    const uint64_t k = blockIdx.y;
    const uint64_t r0 = k * block_rows, r1 = min(rows, r0 + block_rows);
    const uint32_t e0 = row_ptr[r0], e1 = row_ptr[r1];
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (uint64_t i = (uint64_t)e0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e1; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = col_idx[i];
        lo = min(lo, c);
        hi = max(hi, c);
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(out + 2 * k, lo);
        atomicMax(out + 2 * k + 1, hi);
    }
}

int launch_block_col_range(const uint32_t *row_ptr, const uint32_t *col_idx, uint64_t rows, uint64_t block_rows, uint32_t nblocks,
                           uint32_t *out, cudaStream_t stream)
{
    if (nblocks == 0) return BSM_OK;
    if (nblocks > 65535) return fail(BSM_ERR_INVALID_ARGUMENT, "block_col_range: too many row blocks");
    const int gx = std::max(1, grid_for(rows, 256) / (int)nblocks);
    block_col_range_kernel<<<dim3(gx, nblocks), 256, 0, stream>>>(row_ptr, col_idx, rows, block_rows, out);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

// scatter of the zero-dropped entries of a row block straight into the reference's result layout: values and
// usize (u64) column indices, positions from the block-local exclusive scan
template <typename T>
__global__ void scatter_nonzero64_kernel(const T *__restrict__ d, uint64_t rows, uint64_t cols, uint64_t ld,
                                         const uint32_t *__restrict__ row_ptr, T *__restrict__ vals, uint64_t *__restrict__ col_index)
{
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        uint32_t pos = row_ptr[r];
        for (uint64_t c0 = 0; c0 < cols; c0 += 32) {
            const uint64_t c = c0 + lane;
            T x = T(0);
            if (c < cols) x = d[r * ld + c];
            const bool k = c < cols && keep(x);
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, k);
            if (k) {
                const uint32_t o = pos + __popc(m & ((1u << lane) - 1u));   // ascending column = insertion order
                vals[o] = x;
                if (col_index) col_index[o] = c;   // (null: the columns travel as row masks, see count_nonzero_kernel)
            }
            pos += __popc(m);
        }
    }
}

int launch_scatter_nonzero64(int dtype, const void *dense, uint64_t rows, uint64_t cols, uint64_t ld, const uint32_t *row_ptr, void *vals,
                             uint64_t *col_index, cudaStream_t stream)
{
    if (rows == 0) return BSM_OK;
    const int grid = grid_for(rows * 32, 256);
    if (dtype == BSM_F64)
        scatter_nonzero64_kernel<double><<<grid, 256, 0, stream>>>((const double *)dense, rows, cols, ld, row_ptr, (double *)vals, col_index);
    else
        scatter_nonzero64_kernel<float><<<grid, 256, 0, stream>>>((const float *)dense, rows, cols, ld, row_ptr, (float *)vals, col_index);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

// row_index piece of row block k in the reference layout: out64[i] = entries before the block (tot[k]) + the block-local
// exclusive scan; tot[k+1] = tot[k] + entries of the block (local_rp[rows])
// host_tot (mapped pinned host memory) receives tot[k+1] by a plain store over PCIe — NOT by a copy: a cudaMemcpyAsync of these
// 8 bytes queues on the device->host copy engine behind gigabytes of result blocks (measured: ~30 ms late per block)
__global__ void row_index_piece_kernel(const uint32_t *__restrict__ local_rp, uint64_t rows, unsigned long long *__restrict__ tot, uint32_t k,
                                       uint64_t *__restrict__ out64, volatile unsigned long long *host_tot)
{
    const unsigned long long base = tot[k];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += stride) out64[i] = base + local_rp[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long next = base + local_rp[rows];
        tot[k + 1] = next;
        if (host_tot) {
            host_tot[k + 1] = next;
            __threadfence_system();
        }
    }
}

int launch_row_index_piece(const uint32_t *local_rp, uint64_t rows, unsigned long long *tot, uint32_t k, uint64_t *out64,
                           unsigned long long *host_tot, cudaStream_t stream)
{
    row_index_piece_kernel<<<grid_for(std::max<uint64_t>(rows, 1), 256), 256, 0, stream>>>(local_rp, rows, tot, k, out64, host_tot);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

}  // namespace bsm
