// membw.cu — memory-system probe for the SpMM design (not part of the product library).
// Measures on one B200: DRAM read / write / copy bandwidth, L2-resident read bandwidth and the
// gather bandwidth of 128 B .. 1 KB row segments at several working-set sizes. The numbers bound
// what the B-row gather of the SpMM kernels can reach ("memory-system budget").
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membw membw.cu && ./membw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

__global__ void read_kernel(const double2 *__restrict__ p, size_t n, int reps, double *sink)
{
    double acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * 4) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u * stride < n) v[u] = __ldg(p + i + u * stride);
                else v[u] = make_double2(0, 0);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y;
        }
    if (acc == 123.456) *sink = acc;
}

__global__ void write_kernel(double2 *__restrict__ p, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(p + i, make_double2(1.0, 2.0));
}

__global__ void copy_kernel(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * 4) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < n) v[u] = __ldg(a + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < n) __stcs(b + i + u * stride, v[u]);
    }
}

// Gather: each group of `lanes` lanes reads one row segment of lanes*16 bytes at a pseudo-random
// row inside a window of `window_rows` rows; `per` gathers in flight per lane.
template <int PER>
__global__ void gather_kernel(const double2 *__restrict__ p, size_t ld16, uint32_t window_rows, int lanes, size_t iters,
                              double *sink)
{
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / lanes;
    const uint32_t gl = threadIdx.x % lanes;
    uint32_t state = gid * 2654435761u + 12345u;
    double acc = 0;
    for (size_t it = 0; it < iters; ++it) {
        double2 v[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            state = state * 1664525u + 1013904223u;
            const uint32_t row = (uint32_t)(((uint64_t)(state >> 4) * window_rows) >> 28);
            v[u] = __ldg(p + (size_t)row * ld16 + gl);
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) acc += v[u].x + v[u].y;
    }
    if (acc == 123.456) *sink = acc;
}

template <typename F> static float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

// The vector kernel's traversal without its arithmetic: persistent CTAs of `warps` warps; CTA b owns super-batches b, b+grid, ... of
// warps*P consecutive 1 KB rows; warp w copies rows [w*P, (w+1)*P) of it one row (32 lanes x 2 x 16 bytes) at a time with `u` rows
// of loads in flight; `extra` further rows are read per row at the stencil offsets (+-1 line = P rows, +-1 plane) and discarded,
// to reproduce the gather traffic (L1 / L2 hits) next to the streams. What copy bandwidth does the TRAVERSAL allow?
template <int U>
__global__ void __launch_bounds__(256, 3) pattern_copy_kernel(const double2 *__restrict__ a, double2 *__restrict__ b, uint32_t rows, uint32_t P,
                                                              uint32_t plane, int gathers, double *sink)
{
    const uint32_t W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = W * P, num_super = (rows + S - 1) / S;
    double acc = 0.0;
    for (uint32_t sb = blockIdx.x; sb < num_super; sb += gridDim.x) {
        const uint32_t r0 = sb * S + warp * P, r1 = min(rows, r0 + P);
        for (uint32_t r = r0; r < r1; r += U) {
            double2 v[U][2];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r + u < r1) {
                    v[u][0] = __ldg(a + (size_t)(r + u) * 64 + lane);
                    v[u][1] = __ldg(a + (size_t)(r + u) * 64 + 32 + lane);
                }
            if (gathers) {   // the six neighbour rows of a 7-point stencil row (mostly L1 / L2 hits)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (r + u < r1) {
                        const long long offs[6] = {-(long long)plane, -(long long)P, -1, 1, (long long)P, (long long)plane};
#pragma unroll
                        for (int g = 0; g < 6; ++g) {
                            const long long j = (long long)(r + u) + offs[g];
                            if (g < gathers && j >= 0 && j < (long long)rows) {
                                const double2 x = __ldg(a + (size_t)j * 64 + lane), y = __ldg(a + (size_t)j * 64 + 32 + lane);
                                acc += x.x + y.y;
                            }
                        }
                    }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r + u < r1) {
                    __stcs(b + (size_t)(r + u) * 64 + lane, v[u][0]);
                    __stcs(b + (size_t)(r + u) * 64 + 32 + lane, v[u][1]);
                }
        }
    }
    if (acc == 123.456) *sink = acc;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, L2 %.0f MB\n", prop.name, sms, prop.l2CacheSize / 1048576.0);
    const size_t big = (size_t)4 << 30;
    double2 *a, *b;
    double *sink;
    CK(cudaMalloc(&a, big));
    CK(cudaMalloc(&b, big));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(a, 0, big));
    CK(cudaMemset(b, 0, big));
    const size_t n16 = big / 16;
    const int grid = sms * 8, block = 512;

    float ms = time_ms([&] { read_kernel<<<grid, block>>>(a, n16, 1, sink); });
    printf("DRAM read   %8.1f GB/s\n", big / ms / 1e6);
    ms = time_ms([&] { write_kernel<<<grid, block>>>(b, n16); });
    printf("DRAM write  %8.1f GB/s\n", big / ms / 1e6);
    ms = time_ms([&] { copy_kernel<<<grid, block>>>(a, b, n16); });
    printf("DRAM copy   %8.1f GB/s (read+write bytes)\n", 2.0 * big / ms / 1e6);
    ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, big, cudaMemcpyDeviceToDevice)); });
    printf("cudaMemcpy  %8.1f GB/s (read+write bytes)\n", 2.0 * big / ms / 1e6);

    {   // 4 GB = 4 Mi rows of 1 KB (64 planes of a 256^3 grid's B): the vector kernel's traversal as a plain copy
        const uint32_t rows = (uint32_t)(big / 1024);
        for (int gathers : {0, 2, 6})
            for (uint32_t P : {16u, 64u, 256u, 1024u}) {
                ms = time_ms([&] { pattern_copy_kernel<4><<<sms * 3, 256>>>(a, b, rows, P, 65536u, gathers, sink); });
                printf("pattern copy, %d neighbour rows read, 3 CTAs x 8 warps, %4u rows per warp, 4 rows in flight: %8.1f GB/s (read+write bytes of the copy)\n",
                       gathers, P, 2.0 * big / ms / 1e6);
            }
        ms = time_ms([&] { pattern_copy_kernel<8><<<sms * 2, 256>>>(a, b, rows, 256u, 65536u, 0, sink); });
        printf("pattern copy, 0 neighbour rows, 2 CTAs x 8 warps,  256 rows per warp, 8 rows in flight: %8.1f GB/s\n", 2.0 * big / ms / 1e6);
    }

    for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 192}) {
        const size_t bytes = mb << 20;
        const int reps = (int)(((size_t)8 << 30) / bytes);
        ms = time_ms([&] { read_kernel<<<grid, block>>>(a, bytes / 16, reps, sink); });
        printf("resident read  %4zu MB working set: %8.1f GB/s\n", mb, (double)bytes * reps / ms / 1e6);
    }

    // gathers: segment bytes x window bytes
    for (int lanes : {8, 16, 32}) {
        const size_t ld16 = 64;   // 1 KB rows (n = 128 f64)
        for (size_t win_mb : {1, 32, 512, 4096}) {
            const uint32_t window_rows = (uint32_t)((win_mb << 20) / (ld16 * 16));
            const size_t iters = 256;
            const size_t groups = (size_t)grid * block / lanes;
            ms = time_ms([&] { gather_kernel<8><<<grid, block>>>(a, ld16, window_rows, lanes, iters, sink); });
            const double bytes = (double)groups * iters * 8 * lanes * 16;
            printf("gather %4d B segments, window %5zu MB: %8.1f GB/s\n", lanes * 16, win_mb, bytes / ms / 1e6);
        }
    }
    return 0;
}
