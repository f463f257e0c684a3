// spmm_merge.cu — nnz-balanced merge-path SpMM/SpMV for power-law rows (sm_100a).
//
// Same contraction as Csr::mul_dense (/root/reference/src/sparse.rs:431-444), but the work is
// split along the merge path of (row ends, entries) so every lane group gets exactly `items`
// units of (rows closed + entries consumed), whatever the row-length distribution:
//   * merge_partition_kernel: binary search of each chunk boundary's diagonal; depends only on A,
//     computed once per matrix and cached in the handle;
//   * spmm_merge_kernel: a CTA stages the contiguous col_idx/values slice (and the row_ptr slice)
//     of its lane groups' chunks into shared memory with TMA bulk copies (cp.async.bulk), then
//     every group walks its chunk in stored order: gathers B rows with coalesced vector loads,
//     FMA-accumulates, and writes each row it closes exactly once. Entries that belong to a row
//     closed by a later chunk leave the group as one carry-out (row, n partial sums);
//   * merge_fixup_kernel: for every row with carry-outs, adds them in chunk order (deterministic,
//     no atomics): C[row] = (carry_0 + carry_1 + ...) + tail.
// Long rows are therefore summed as a few partial sums instead of one left-to-right chain —
// within the stated tolerance of the reference (and exact for exactly representable data).
#include <algorithm>

#include "bsm_common.cuh"
#include "kernels.h"
#include "spmm_stream.cuh"

namespace bsm {

constexpr uint32_t kNoCarry = 0xFFFFFFFFu;

// One thread per chunk boundary: first row the chunk closes (CUB-style merge-path search over
// list A = row_end[i] = row_ptr[i+1] and list B = 0..nnz-1).
__global__ void merge_partition_kernel(const uint32_t *__restrict__ row_ptr, uint32_t rows, uint32_t nnz, uint32_t items,
                                       uint32_t num_chunks, uint32_t *__restrict__ part_rows)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > num_chunks) return;
    const uint64_t total = (uint64_t)rows + nnz;
    const uint64_t d64 = min((uint64_t)c * items, total);
    const uint32_t d = (uint32_t)d64;
    uint32_t lo = d > nnz ? d - nnz : 0u;
    uint32_t hi = min(d, rows);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(row_ptr + mid + 1) <= d - mid - 1)
            lo = mid + 1;
        else
            hi = mid;
    }
    part_rows[c] = lo;
}

constexpr uint32_t kMergeSlack = 40;   // staged arrays are padded: the gather engine's LDS.128 run past the chunk

template <typename T, int V, int G, int NT, bool FULLN, int U>
__global__ void __launch_bounds__(256) spmm_merge_kernel(const MergeParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;

    const uint32_t groups_per_cta = (blockDim.x / 32) * (32 / G);
    const uint32_t span = groups_per_cta * p.items;   // merge items per CTA

    // shared layout: vals[span+slack] | idx[span+slack] | rp[span+8]
    T *val_s = reinterpret_cast<T *>(smem);
    uint32_t *idx_s = reinterpret_cast<uint32_t *>(smem + (size_t)(span + kMergeSlack) * sizeof(T));
    uint32_t *rp_s = idx_s + (span + kMergeSlack);

    const uint32_t c0 = blockIdx.x * groups_per_cta;
    const uint32_t cend = min(c0 + groups_per_cta, p.num_chunks);
    const uint32_t total = p.rows + p.nnz;
    const uint32_t R0 = __ldg(p.part_rows + c0);
    const uint32_t R1 = __ldg(p.part_rows + cend);
    const uint32_t Z0 = c0 * p.items - R0;
    const uint32_t Z1 = min(cend * p.items, total) - R1;
    const uint32_t z_a = Z0 & ~3u;                       // aligned first staged entry
    const uint32_t rp_a = (R0 + 1u) & ~3u;               // aligned first staged row_ptr index
    const uint32_t cnt = Z1 > Z0 ? ((Z1 - z_a + 3u) & ~3u) : 0u;
    const uint32_t cnt_r = R1 > R0 ? ((R1 + 1u - rp_a + 3u) & ~3u) : 0u;   // rp[R0+1 .. R1]

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t policy = (p.flags & BSM_TUNE_A_EVICT_FIRST) ? l2_policy_evict_first() : l2_policy_evict_normal();
        mbar_arrive_expect_tx(&bar, cnt * (4u + (uint32_t)sizeof(T)) + cnt_r * 4u);
        if (cnt) {
            bulk_g2s(idx_s, p.col_idx + z_a, cnt * 4u, &bar, policy);
            bulk_g2s(val_s, static_cast<const T *>(p.vals) + z_a, cnt * (uint32_t)sizeof(T), &bar, policy);
        }
        if (cnt_r) bulk_g2s(rp_s, p.row_ptr + rp_a, cnt_r * 4u, &bar, policy);
    }

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t grp = (threadIdx.x >> 5) * (32 / G) + lane / G;
    const uint32_t gl = lane % G;
    const uint32_t c = c0 + grp;
    const bool active = c < cend;

    uint32_t row = 0, row_next = 0, nz = 0, nz_end = 0;
    if (active) {
        row = __ldg(p.part_rows + c);
        row_next = __ldg(p.part_rows + c + 1);
        nz = c * p.items - row;
        nz_end = min((c + 1) * p.items, total) - row_next;
    }
    bool col_ok[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) col_ok[t] = FULLN || (uint32_t)((t * G + gl) * V) < p.n;
    const char *__restrict__ b_bytes = reinterpret_cast<const char *>(static_cast<const T *>(p.B) + gl * V);
    const uint32_t ldb_bytes = p.ldb * (uint32_t)sizeof(T);
    const uint32_t ldc_bytes = p.ldc * (uint32_t)sizeof(T);
    const bool streaming = (p.flags & BSM_TUNE_C_STREAMING) != 0;

    mbar_wait(&bar, 0);   // staged slices have landed
    if (!active) return;

    Lane<T, V> acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t].zero();
    bool dirty = false;   // entries accumulated since the last row close
    const uint32_t *rp_rel = rp_s - rp_a;   // row_ptr[i] at rp_rel[i]
    uint32_t row_end = row < row_next ? rp_rel[row + 1] : kNoCarry;
    char *crow = reinterpret_cast<char *>(static_cast<T *>(p.C) + gl * V) + (size_t)row * ldc_bytes;

    auto close_row = [&]() {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (FULLN || col_ok[t]) acc[t].store(reinterpret_cast<T *>(crow) + t * G * V, streaming);
            acc[t].zero();
        }
        crow += ldc_bytes;
        dirty = false;
        ++row;
        row_end = row < row_next ? rp_rel[row + 1] : kNoCarry;
    };

    // stored order inside the chunk, FMA; rows (possibly empty) that end before an entry are closed first
    // LDS.128 reads of the staged A stream (same-box A/B: 3.17 vs 4.16 ms scalar on R-MAT f64) and evict_last
    // on the B gathers (3.17 -> 3.07 ms): profiles/r1_ab2_rmat_f64_areads.jsonl, r1_ab3_evict_last.jsonl
    stream_entries<T, V, NT, FULLN, U, true, true, false, true>(idx_s - z_a, val_s - z_a, nz, nz_end, b_bytes, ldb_bytes, col_ok, G, acc,
                                                   [&](uint32_t k) {
                                                       while (k >= row_end) close_row();
                                                       dirty = true;
                                                   });
    while (row < row_next) close_row();   // rows ending exactly at the chunk end, trailing empty rows

    // whatever is left belongs to row_next, which a later chunk closes
    if (dirty) {
        T *car = static_cast<T *>(p.carry_vals) + (size_t)c * p.ldcar + gl * V;
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if (FULLN || col_ok[t]) acc[t].store(car + t * G * V, false);
    }
}

// ---- fix-up: C[row] = (carry_first + ... + carry_last) + tail ---------------------------------------
// Which chunks carry into a row follows from row_ptr alone: entry e of row r is merge item e + r and
// the row's end is item row_ptr[r+1] + r, so chunks [ (row_ptr[r]+r)/items , (row_ptr[r+1]+r)/items )
// hold entries of r without closing it (each of them left one carry-out), and the chunk that closes it
// wrote the tail. Carries are added in chunk order, before the tail (they precede it in stored order).
constexpr uint32_t kFixupSerialMax = 64;   // longer runs go to the cooperative kernel

// One warp per row. Short runs are summed here; long runs (hub rows of a power-law matrix, thousands of
// chunks) are queued for merge_fixup_long_kernel.
template <typename T>
__global__ void __launch_bounds__(256) merge_fixup_kernel(const MergeParams p)
{
    // (warp index, not thread index: rows * 32 exceeds 2^32 above 134 M rows)
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (row >= p.rows) return;
    const uint32_t c = (__ldg(p.row_ptr + row) + row) / p.items;
    const uint32_t c_close = (__ldg(p.row_ptr + row + 1) + row) / p.items;
    if (c_close <= c) return;
    const uint32_t len = c_close - c;
    if (len > kFixupSerialMax) {
        if (lane == 0) {
            const uint32_t slot = atomicAdd(p.long_count, 1u);
            if (slot < p.long_cap) {
                p.long_rows[2 * slot] = row;
                p.long_rows[2 * slot + 1] = c;
            }
        }
        return;
    }
    const T *car = static_cast<const T *>(p.carry_vals);
    T *crow = static_cast<T *>(p.C) + (size_t)row * p.ldc;
    for (uint32_t j = lane; j < p.n; j += 32) {
        T sum = car[(size_t)c * p.ldcar + j];
#pragma unroll 8
        for (uint32_t k = 1; k < len; ++k) sum += car[(size_t)(c + k) * p.ldcar + j];
        crow[j] = sum + crow[j];
    }
}

// One CTA (8 warps) per long run: warp w sums the w-th of 8 contiguous segments of the run in chunk
// order, then the 8 partial sums are combined in segment order — deterministic for a given matrix.
template <typename T>
__global__ void __launch_bounds__(256) merge_fixup_long_kernel(const MergeParams p)
{
    constexpr uint32_t kMaxN = 4096 / sizeof(T);   // columns of one pass (<= 32 lanes x 16 bytes x 4 tiles)
    __shared__ T partial[8][kMaxN];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t count = min(*p.long_count, p.long_cap);
    const T *car = static_cast<const T *>(p.carry_vals);
    for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
        const uint32_t row = p.long_rows[2 * i], c = p.long_rows[2 * i + 1];
        const uint32_t len = (__ldg(p.row_ptr + row + 1) + row) / p.items - c;
        const uint32_t seg = (len + 7) / 8;
        const uint32_t k0 = min(warp * seg, len), k1 = min(k0 + seg, len);
        for (uint32_t j = lane; j < p.n; j += 32) {
            T sum = T(0);
#pragma unroll 8
            for (uint32_t k = k0; k < k1; ++k) sum += car[(size_t)(c + k) * p.ldcar + j];
            partial[warp][j] = sum;
        }
        __syncthreads();
        T *crow = static_cast<T *>(p.C) + (size_t)row * p.ldc;
        for (uint32_t j = threadIdx.x; j < p.n; j += blockDim.x) {
            T sum = partial[0][j];
#pragma unroll
            for (int w = 1; w < 8; ++w) sum += partial[w][j];
            crow[j] = sum + crow[j];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
constexpr int merge_default_u(int NT) { return NT >= 4 ? 4 : (NT == 2 ? 4 : 8); }

// The unpredicated (FULLN) form is built for the shapes that carry the measured workloads — 128-bit lanes, a full warp with one
// or two register tiles, and the grouped shapes; every other shape runs the predicated form, whose extra cost disappears behind
// the gathers (the kernel is bound by the L1 data pipe and L2, §3.2 of DESIGN.md). Halves the size of this object.
template <typename T, int V, int G, int NT> static const void *merge_kernel_ptr(bool fulln)
{
    constexpr int U = merge_default_u(NT);
    constexpr bool kBuildFull = V * sizeof(T) == 16 && ((G == 32 && NT <= 2) || (G < 32 && NT == 2));
    if constexpr (kBuildFull) {
        if (fulln) return reinterpret_cast<const void *>(&spmm_merge_kernel<T, V, G, NT, true, U>);
    }
    if constexpr (G < 32 && NT > 1)
        return nullptr;   // the grouped shapes exist for full-width rows only
    else
        return reinterpret_cast<const void *>(&spmm_merge_kernel<T, V, G, NT, false, U>);
}
template <typename T, int V> static const void *merge_kernel_select_gnt(int G, int NT, bool fulln)
{
    if (G == 32) {
        switch (NT) {
            case 1: return merge_kernel_ptr<T, V, 32, 1>(fulln);
            case 2: return merge_kernel_ptr<T, V, 32, 2>(fulln);
            case 4: return merge_kernel_ptr<T, V, 32, 4>(fulln);
        }
        return nullptr;
    }
    if (NT != 1) {
        // several register tiles per lane on fewer lanes: 32/G chunks side by side per warp, so one LDS of the
        // staged A stream feeds 32/G entries (128-bit lanes, full-width shapes only)
        if constexpr (V * sizeof(T) == 16) {
            if (!fulln) return nullptr;
            if (NT == 2) {   // the grouped shapes the sweeps kept (profiles/r1_sweep{x,y}_rmat_*.jsonl); 4 tiles never won
                switch (G) {
                    case 16: return merge_kernel_ptr<T, V, 16, 2>(true);
                    case 8: return merge_kernel_ptr<T, V, 8, 2>(true);
                }
            }
        }
        return nullptr;
    }
    switch (G) {
        case 16: return merge_kernel_ptr<T, V, 16, 1>(fulln);
        case 8: return merge_kernel_ptr<T, V, 8, 1>(fulln);
        case 4: return merge_kernel_ptr<T, V, 4, 1>(fulln);
        case 2: return merge_kernel_ptr<T, V, 2, 1>(fulln);
        case 1: return merge_kernel_ptr<T, V, 1, 1>(fulln);
    }
    return nullptr;
}
static const void *merge_kernel_select(int dtype, Shape sh, uint32_t n)
{
    const bool fulln = n == (uint32_t)(sh.V * sh.G * sh.NT);
    if (dtype == BSM_F64) {
        if (sh.V == 1) return merge_kernel_select_gnt<double, 1>(sh.G, sh.NT, fulln);
        if (sh.V == 2) return merge_kernel_select_gnt<double, 2>(sh.G, sh.NT, fulln);
    } else {
        if (sh.V == 1) return merge_kernel_select_gnt<float, 1>(sh.G, sh.NT, fulln);
        if (sh.V == 4) return merge_kernel_select_gnt<float, 4>(sh.G, sh.NT, fulln);
    }
    return nullptr;
}

size_t merge_kernel_smem_bytes(int dtype, Shape sh, int block, uint32_t items)
{
    const size_t span = (size_t)(block / 32) * (32 / sh.G) * items;
    return (span + kMergeSlack) * dtype_size(dtype) + (span + kMergeSlack) * 4 + (span + 8) * 4;
}

int launch_merge_partition(const uint32_t *row_ptr, uint32_t rows, uint32_t nnz, uint32_t items, uint32_t num_chunks,
                           uint32_t *part_rows, cudaStream_t stream)
{
    const uint32_t threads = 256;
    const uint32_t blocks = (num_chunks + 1 + threads - 1) / threads;
    merge_partition_kernel<<<blocks, threads, 0, stream>>>(row_ptr, rows, nnz, items, num_chunks, part_rows);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

int launch_spmm_merge(int dtype, Shape sh, const MergeParams &p, int block, size_t smem, int ctas_per_sm, cudaStream_t stream, int *grid_out)
{
    const void *k = merge_kernel_select(dtype, sh, p.n);
    if (!k) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_merge: no kernel for this lane shape");
    const uint32_t groups_per_cta = (uint32_t)(block / 32) * (32 / sh.G);
    const uint32_t grid = (p.num_chunks + groups_per_cta - 1) / groups_per_cta;
    if (grid == 0) return BSM_OK;
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Shared-memory carve-out: exactly what the resident CTAs' stages need — the rest of the 228 KB stays L1, where the hot B
    // rows of a power-law matrix live. ctas_per_sm > 0 caps the residency below what registers allow (fewer warps, larger L1).
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int occ = 0;
    BSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, block, smem));
    if (occ < 1) return fail(BSM_ERR_CUDA, "spmm_merge: kernel does not fit on an SM");
    const int resident = ctas_per_sm > 0 ? std::min(ctas_per_sm, occ) : occ;
    const size_t need = (size_t)resident * (smem + 1024 + 64);   // + the driver's per-CTA reservation and the static barrier
    const int pct = (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024));
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    MergeParams pc = p;
    void *args[] = {&pc};
    BSM_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(block), args, smem, stream));
    count_launch();
    if (grid_out) *grid_out = (int)grid;
    return BSM_OK;
}

int launch_merge_fixup(int dtype, const MergeParams &p, cudaStream_t stream, int *launched)
{
    if (launched) *launched = 0;
    if (p.num_chunks == 0 || p.rows == 0) return BSM_OK;
    const uint32_t threads = 256;
    const uint64_t total_threads = (uint64_t)p.rows * 32;
    const uint32_t blocks = (uint32_t)((total_threads + threads - 1) / threads);
    BSM_CUDA(cudaMemsetAsync(p.long_count, 0, 4, stream));
    if (dtype == BSM_F64)
        merge_fixup_kernel<double><<<blocks, threads, 0, stream>>>(p);
    else
        merge_fixup_kernel<float><<<blocks, threads, 0, stream>>>(p);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    if (launched) ++*launched;
    if (p.long_cap) {   // a matrix can only have runs longer than kFixupSerialMax chunks if it is that large
        const uint32_t grid = std::min<uint32_t>(p.long_cap, 592);
        if (dtype == BSM_F64)
            merge_fixup_long_kernel<double><<<grid, threads, 0, stream>>>(p);
        else
            merge_fixup_long_kernel<float><<<grid, threads, 0, stream>>>(p);
        BSM_CUDA(cudaGetLastError());
        count_launch();
        if (launched) ++*launched;
    }
    return BSM_OK;
}

}  // namespace bsm
