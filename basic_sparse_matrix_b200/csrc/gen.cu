// gen.cu — synthetic workloads generated directly in HBM (bench / test support, not the hot path).
//
// Everything is a pure function of (seed, logical index) through the counter-based hash below,
// so the CPU side (basic_sparse_matrix_b200/gen.py, numpy) regenerates any element independently
// and the oracle can check sampled rows of full-size workloads without moving them over PCIe.
// The reference's own bench draws from rand 0.8.5's StdRng (ChaCha12;
// /root/reference/benches/sparse_dense_mul.rs:16), which cannot be reproduced-and-verified
// without a Rust toolchain — the shapes and densities are the bench's, the streams are ours.
#include <cub/device/device_radix_sort.cuh>

#include "bsm_common.cuh"
#include "kernels.h"

namespace bsm {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// h(seed, i) — mirrored by gen.py:hash_u64
__host__ __device__ __forceinline__ uint64_t hash_u64(uint64_t seed, uint64_t i)
{
    return mix64((i + 1ULL) * 0x9E3779B97F4A7C15ULL + seed * 0xD1342543DE82EF95ULL);
}
// mode 0: k/1024, k in [0,1024) ; mode 2: k/16, k in [0,16) ; mode 1: uniform [0,1); all + offset
__host__ __device__ __forceinline__ double dense_value(uint64_t h, int mode, double offset)
{
    const uint64_t m = h >> 11;
    if (mode == 0) return (double)(m % 1024ULL) / 1024.0 + offset;
    if (mode == 2) return (double)(m % 16ULL) / 16.0 + offset;
    return (double)m * (1.0 / 9007199254740992.0) + offset;
}
// matrix values: mode 0: k/256, k in [1,256] ; mode 2: k/8, k in [1,8] ; mode 1: uniform [0.5,1.5)
__host__ __device__ __forceinline__ double sparse_value(uint64_t h, int mode)
{
    const uint64_t m = h >> 11;
    if (mode == 0) return (double)(1ULL + m % 256ULL) / 256.0;
    if (mode == 2) return (double)(1ULL + m % 8ULL) / 8.0;
    return 0.5 + (double)m * (1.0 / 9007199254740992.0);
}

template <typename T>
__global__ void gen_dense_kernel(T *data, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t seed, int mode, double offset)
{
    const uint64_t total = rows * cols;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint64_t r = i / cols, c = i % cols;
        data[r * ld + c] = (T)dense_value(hash_u64(seed, i), mode, offset);
    }
}

int launch_gen_dense(int dtype, void *data, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t seed, int mode,
                     double offset, cudaStream_t stream)
{
    if (rows * cols == 0) return BSM_OK;
    const int grid = 148 * 16;
    if (dtype == BSM_F64)
        gen_dense_kernel<double><<<grid, 256, 0, stream>>>((double *)data, rows, cols, ld, seed, mode, offset);
    else
        gen_dense_kernel<float><<<grid, 256, 0, stream>>>((float *)data, rows, cols, ld, seed, mode, offset);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

// ---- Laplacians --------------------------------------------------------------------------------
__global__ void laplacian_counts_kernel(uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_begin, uint64_t row_end,
                                        uint32_t *counts)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < row_end - row_begin; k += stride) {
        const uint64_t i = row_begin + k;
        const uint64_t x = i % nx, y = (i / nx) % ny, z = i / (nx * ny);
        counts[k] = 1u + (x > 0) + (x + 1 < nx) + (y > 0) + (y + 1 < ny) + (z > 0) + (z + 1 < nz);
    }
}

template <typename T>
__global__ void laplacian_fill_kernel(uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_begin, uint64_t row_end,
                                      const uint32_t *__restrict__ row_ptr, uint32_t *__restrict__ col_idx,
                                      T *__restrict__ vals)
{
    const T diag = (T)(2.0 * (double)((nx > 1) + (ny > 1) + (nz > 1)));
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < row_end - row_begin; k += stride) {
        const uint64_t i = row_begin + k;
        const uint64_t x = i % nx, y = (i / nx) % ny, z = i / (nx * ny);
        uint32_t o = row_ptr[k];
        // columns ascending: z-1, y-1, x-1, self, x+1, y+1, z+1
        if (z > 0) { col_idx[o] = (uint32_t)(i - nx * ny); vals[o++] = (T)-1; }
        if (y > 0) { col_idx[o] = (uint32_t)(i - nx); vals[o++] = (T)-1; }
        if (x > 0) { col_idx[o] = (uint32_t)(i - 1); vals[o++] = (T)-1; }
        col_idx[o] = (uint32_t)i; vals[o++] = diag;
        if (x + 1 < nx) { col_idx[o] = (uint32_t)(i + 1); vals[o++] = (T)-1; }
        if (y + 1 < ny) { col_idx[o] = (uint32_t)(i + nx); vals[o++] = (T)-1; }
        if (z + 1 < nz) { col_idx[o] = (uint32_t)(i + nx * ny); vals[o++] = (T)-1; }
    }
}

int launch_laplacian_counts(uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_begin, uint64_t row_end, uint32_t *counts,
                            cudaStream_t stream)
{
    if (row_end <= row_begin) return BSM_OK;
    laplacian_counts_kernel<<<148 * 8, 256, 0, stream>>>(nx, ny, nz, row_begin, row_end, counts);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

int launch_laplacian_fill(int dtype, uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_begin, uint64_t row_end,
                          const uint32_t *row_ptr, uint32_t *col_idx, void *vals, cudaStream_t stream)
{
    if (row_end <= row_begin) return BSM_OK;
    if (dtype == BSM_F64)
        laplacian_fill_kernel<double><<<148 * 8, 256, 0, stream>>>(nx, ny, nz, row_begin, row_end, row_ptr, col_idx, (double *)vals);
    else
        laplacian_fill_kernel<float><<<148 * 8, 256, 0, stream>>>(nx, ny, nz, row_begin, row_end, row_ptr, col_idx, (float *)vals);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

// ---- SPD band ------------------------------------------------------------------------------------
__global__ void band_counts_kernel(uint64_t n, uint64_t hb, uint64_t row_begin, uint64_t row_end, uint32_t *counts)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < row_end - row_begin; k += stride) {
        const uint64_t i = row_begin + k;
        const uint64_t lo = i < hb ? i : hb, hi = (n - 1 - i) < hb ? (n - 1 - i) : hb;
        counts[k] = (uint32_t)(lo + 1 + hi);
    }
}

template <typename T>
__global__ void band_fill_kernel(uint64_t n, uint64_t hb, uint64_t row_begin, uint64_t row_end,
                                 const uint32_t *__restrict__ row_ptr, uint32_t *__restrict__ col_idx, T *__restrict__ vals)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < row_end - row_begin; k += stride) {
        const uint64_t i = row_begin + k;
        const uint64_t lo = i < hb ? i : hb, hi = (n - 1 - i) < hb ? (n - 1 - i) : hb;
        uint32_t o = row_ptr[k];
        double diag = 1.0;   // a_ii = 1 + sum_j |a_ij|, summed over ascending j in f64
        for (uint64_t d = lo; d >= 1; --d) {
            const double a = 1.0 / (1.0 + (double)d);
            col_idx[o] = (uint32_t)(i - d);
            vals[o++] = (T)(-a);
            diag += a;
        }
        const uint32_t odiag = o++;
        for (uint64_t d = 1; d <= hi; ++d) {
            const double a = 1.0 / (1.0 + (double)d);
            col_idx[o] = (uint32_t)(i + d);
            vals[o++] = (T)(-a);
            diag += a;
        }
        col_idx[odiag] = (uint32_t)i;
        vals[odiag] = (T)diag;
    }
}

int launch_band_counts(uint64_t n, uint64_t hb, uint64_t row_begin, uint64_t row_end, uint32_t *counts, cudaStream_t stream)
{
    if (row_end <= row_begin) return BSM_OK;
    band_counts_kernel<<<148 * 8, 256, 0, stream>>>(n, hb, row_begin, row_end, counts);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

int launch_band_fill(int dtype, uint64_t n, uint64_t hb, uint64_t row_begin, uint64_t row_end, const uint32_t *row_ptr,
                     uint32_t *col_idx, void *vals, cudaStream_t stream)
{
    if (row_end <= row_begin) return BSM_OK;
    if (dtype == BSM_F64)
        band_fill_kernel<double><<<148 * 8, 256, 0, stream>>>(n, hb, row_begin, row_end, row_ptr, col_idx, (double *)vals);
    else
        band_fill_kernel<float><<<148 * 8, 256, 0, stream>>>(n, hb, row_begin, row_end, row_ptr, col_idx, (float *)vals);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

// ---- R-MAT -------------------------------------------------------------------------------------
__global__ void rmat_edges_kernel(int scale, uint64_t edges, double a, double b, double c, uint64_t seed, uint64_t *keys)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const double ab = a + b, abc = a + b + c;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < edges; e += stride) {
        uint64_t row = 0, col = 0;
        for (int l = 0; l < scale; ++l) {
            const double u = (double)(hash_u64(seed, e * 64ULL + (uint64_t)l) >> 11) * (1.0 / 9007199254740992.0);
            const uint64_t rbit = u >= ab ? 1u : 0u;
            const uint64_t cbit = (u >= a && u < ab) || u >= abc ? 1u : 0u;
            row = (row << 1) | rbit;
            col = (col << 1) | cbit;
        }
        keys[e] = (row << 32) | col;
    }
}

template <typename T>
__global__ void rmat_finish_kernel(const uint64_t *__restrict__ keys, uint64_t edges, uint64_t seed, int mode,
                                   uint32_t *__restrict__ col_idx, T *__restrict__ vals)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < edges; e += stride) {
        col_idx[e] = (uint32_t)(keys[e] & 0xFFFFFFFFULL);
        vals[e] = (T)sparse_value(hash_u64(seed + 1ULL, e), mode);
    }
}

// row_ptr[r] = number of sorted keys with row < r  (lower bound of r<<32)
__global__ void rmat_rowptr_kernel(const uint64_t *__restrict__ keys, uint64_t edges, uint64_t rows, uint32_t *row_ptr)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= rows; r += stride) {
        const uint64_t target = r << 32;
        uint64_t lo = 0, hi = edges;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (keys[mid] < target)
                lo = mid + 1;
            else
                hi = mid;
        }
        row_ptr[r] = (uint32_t)lo;
    }
}

int gen_rmat_device(int dtype, int scale, uint64_t edges, double a, double b, double c, uint64_t seed, int mode,
                    uint32_t *row_ptr, uint32_t *col_idx, void *vals, cudaStream_t stream)
{
    if (scale < 1 || scale > 31) return fail(BSM_ERR_INVALID_ARGUMENT, "rmat: scale must be in [1,31]");
    const uint64_t rows = 1ULL << scale;
    uint64_t *keys_in = nullptr, *keys_out = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    BSM_CUDA(cudaMallocAsync(&keys_in, (edges ? edges : 1) * sizeof(uint64_t), stream));
    BSM_CUDA(cudaMallocAsync(&keys_out, (edges ? edges : 1) * sizeof(uint64_t), stream));
    if (edges) {
        rmat_edges_kernel<<<148 * 16, 256, 0, stream>>>(scale, edges, a, b, c, seed, keys_in);
        BSM_CUDA(cudaGetLastError());
        count_launch();
        // library sort for data preparation only (not on the SpMM path)
        BSM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys_in, keys_out, edges, 0, 32 + scale, stream));
        BSM_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, stream));
        BSM_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys_in, keys_out, edges, 0, 32 + scale, stream));
        if (dtype == BSM_F64)
            rmat_finish_kernel<double><<<148 * 16, 256, 0, stream>>>(keys_out, edges, seed, mode, col_idx, (double *)vals);
        else
            rmat_finish_kernel<float><<<148 * 16, 256, 0, stream>>>(keys_out, edges, seed, mode, col_idx, (float *)vals);
        BSM_CUDA(cudaGetLastError());
        count_launch();
    }
    rmat_rowptr_kernel<<<148 * 8, 256, 0, stream>>>(keys_out, edges, rows, row_ptr);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    if (tmp) BSM_CUDA(cudaFreeAsync(tmp, stream));
    BSM_CUDA(cudaFreeAsync(keys_in, stream));
    BSM_CUDA(cudaFreeAsync(keys_out, stream));
    return BSM_OK;
}

}  // namespace bsm
