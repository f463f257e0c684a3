// spmm_rows.cu — vector-CSR SpMM/SpMV for short and regular rows (sm_100a).
//
// Replaces the loop nest of Csr::mul_dense, /root/reference/src/sparse.rs:431-444:
//     for row { row = get_row_compact(row); for out_col { value = 0;
//         for entry in row (stored order) { value = value + entry.v * B[entry.col][out_col] } } }
//
// Design (B200):
//   * persistent CTAs of W warps. A CTA owns "super-batches" of W*P consecutive rows
//     (blockIdx, +grid, ...); inside one, warp w owns the P consecutive rows [w*P, (w+1)*P) and
//     walks them in slices of R rows. With P = one grid line of a stencil matrix the W warps of a
//     CTA sweep W adjacent lines side by side, so the +-1-line neighbours of B are L1 hits, not
//     L2 traffic (optionally kept in step by a CTA barrier every few rows);
//   * every warp is its own TMA pipeline: lane 0 streams the slice's contiguous piece of
//     col_idx / values and its row_ptr window into a warp-private shared-memory ring with bulk
//     copies (cp.async.bulk -> UBLKCP) that complete on per-stage mbarriers, `stages-1` slices
//     ahead of the slice being consumed. The A stream never occupies registers or L1;
//   * a group of G lanes owns one output row; every lane owns V consecutive columns per register
//     tile (V*sizeof(T) up to 16 bytes -> 128-bit coalesced B-row loads on the read-only path).
//     G == 32: the warp treats its slice as ONE flat entry stream — U B-row gathers are always in
//     flight, whatever the row lengths — and closes a row (one streaming store of C) whenever the
//     stream crosses a row end. G < 32: 32/G rows side by side, row by row;
//   * entries are consumed IN STORED ORDER with a separately rounded multiply and add
//     -> bit-identical to the reference's sequential sum for any input.
//
// A slice whose entry count exceeds the stage capacity (irregular matrices forced onto this
// kernel) reads col_idx/values straight from global memory — same arithmetic.
#include "bsm_common.cuh"
#include "kernels.h"

namespace bsm {

constexpr int kMaxStages = 8;
constexpr int kSyncBarrierId = 1;

struct RowSmemLayout {
    uint32_t vals_off, idx_off, rp_off, stage_bytes;
};
__host__ __device__ inline RowSmemLayout row_layout(uint32_t cap, uint32_t R, uint32_t tsize)
{
    RowSmemLayout l;
    l.vals_off = 0;
    l.idx_off = cap * tsize;                  // cap % 4 == 0 -> 16-byte aligned
    l.rp_off = l.idx_off + cap * 4;
    l.stage_bytes = l.rp_off + (R + 4) * 4;   // R % 4 == 0
    return l;
}

size_t row_kernel_smem_bytes(int dtype, const RowParams &p, int warps)
{
    const size_t ring = (size_t)row_layout(p.cap, p.R, (uint32_t)dtype_size(dtype)).stage_bytes * p.stages * warps;
    return ring + (size_t)warps * p.stages * 8;   // + one mbarrier per (warp, stage)
}

__device__ __forceinline__ void cta_bar_sync(uint32_t threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(kSyncBarrierId), "r"(threads) : "memory");
}

template <typename T, int V, int NT, bool FULLN>
__device__ __forceinline__ void load_brow(Lane<T, V> (&b)[NT], const T *__restrict__ brow, const bool (&col_ok)[NT], int G)
{
#pragma unroll
    for (int t = 0; t < NT; ++t)
        if (FULLN || col_ok[t]) b[t].load(brow + t * G * V, false);
}

template <typename T, int V, int NT>
__device__ __forceinline__ void fma_row(Lane<T, V> (&acc)[NT], const Lane<T, V> (&b)[NT], T a)
{
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[t].x[i] = mul_add<false>(a, b[t].x[i], acc[t].x[i]);   // sparse.rs:438-439
}

template <typename T, int V, int G, int NT, bool FULLN>
__global__ void __launch_bounds__(512) spmm_rows_kernel(const RowParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];

    constexpr int U = NT >= 4 ? 2 : (NT == 2 ? 4 : 8);            // B-row gathers in flight per lane group
    constexpr int RPP = 32 / G;                                // rows side by side in one warp (G < 32)

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t W = blockDim.x >> 5;
    const RowSmemLayout L = row_layout(p.cap, p.R, sizeof(T));
    unsigned char *ring = smem + (size_t)warp * p.stages * L.stage_bytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)W * p.stages * L.stage_bytes) + warp * p.stages;

    if (lane == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    const T *__restrict__ vals = static_cast<const T *>(p.vals);
    const uint32_t S = W * p.P;                 // rows per super-batch
    const uint32_t spw = p.P / p.R;             // slices per warp per super-batch
    const uint32_t my_supers = blockIdx.x < p.num_super ? (p.num_super - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const uint32_t my_slices = my_supers * spw;

    // first row of this warp's i-th slice
    auto slice_row0 = [&](uint32_t i) -> uint64_t {
        const uint32_t k = i / spw, t = i - k * spw;
        return (uint64_t)(blockIdx.x + k * gridDim.x) * S + (uint64_t)warp * p.P + (uint64_t)t * p.R;
    };

    // ---- producer side (lane 0): TMA bulk copies of one slice into ring stage i % stages --------
    const uint64_t policy = (p.flags & BSM_TUNE_A_EVICT_FIRST) ? l2_policy_evict_first() : l2_policy_evict_normal();
    uint32_t pf_s = 0, pf_e = 0;                // entry range of the next slice to issue (prefetched)
    auto prefetch_bounds = [&](uint32_t i) {
        if (i < my_slices) {
            const uint64_t r0 = slice_row0(i);
            if (r0 < p.rows) {
                const uint32_t r1 = (uint32_t)min(r0 + p.R, (uint64_t)p.rows);
                pf_s = __ldg(p.row_ptr + r0);
                pf_e = __ldg(p.row_ptr + r1);
            }
        }
    };
    auto issue = [&](uint32_t i) {
        // only lane 0 calls this
        const uint64_t r0 = slice_row0(i);
        if (r0 < p.rows) {
            const uint32_t nr = (uint32_t)min((uint64_t)p.R, p.rows - r0);
            const uint32_t stage = i % p.stages;
            unsigned char *st = ring + (size_t)stage * L.stage_bytes;
            const uint32_t base = pf_s & ~3u;                     // 16-byte aligned start for u32 and T
            const uint32_t cnt = (pf_e - base + 3u) & ~3u;        // entries, multiple of 4
            const uint32_t cnt_r = (nr + 1u + 3u) & ~3u;          // row_ptr window rp[r0 .. r0+nr]
            const bool fast = cnt <= p.cap;
            const uint32_t bytes = cnt_r * 4u + ((fast && cnt) ? cnt * (4u + (uint32_t)sizeof(T)) : 0u);
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            bulk_g2s(st + L.rp_off, p.row_ptr + r0, cnt_r * 4u, &full_bar[stage], policy);
            if (fast && cnt) {
                bulk_g2s(st + L.idx_off, p.col_idx + base, cnt * 4u, &full_bar[stage], policy);
                bulk_g2s(st + L.vals_off, vals + base, cnt * (uint32_t)sizeof(T), &full_bar[stage], policy);
            }
        }
        prefetch_bounds(i + 1);
    };

    if (lane == 0) {
        prefetch_bounds(0);
        for (uint32_t i = 0; i + 1 < p.stages && i < my_slices; ++i) issue(i);
    }

    // ---- consumer side -------------------------------------------------------------------------
    const uint32_t grp = lane / G;    // which of the warp's concurrent rows (G < 32)
    const uint32_t gl = lane % G;     // lane inside the group
    bool col_ok[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) col_ok[t] = FULLN || (uint32_t)((t * G + gl) * V) < p.n;
    const T *__restrict__ b_lane = static_cast<const T *>(p.B) + gl * V;
    T *__restrict__ c_lane = static_cast<T *>(p.C) + gl * V;
    const bool streaming = (p.flags & BSM_TUNE_C_STREAMING) != 0;
    const uint32_t sync_rows = p.sync_rows;
    const uint32_t cta_threads = blockDim.x;

    for (uint32_t i = 0; i < my_slices; ++i) {
        __syncwarp();   // every lane is done reading the stage that is refilled next
        if (lane == 0 && i + p.stages - 1 < my_slices) issue(i + p.stages - 1);

        const uint64_t row0_64 = slice_row0(i);
        uint32_t barriers_left = sync_rows ? p.R / sync_rows : 0u;
        if (row0_64 < p.rows) {
            const uint32_t row0 = (uint32_t)row0_64;
            const uint32_t nr = min(p.R, p.rows - row0);
            const uint32_t stage = i % p.stages;
            mbar_wait(&full_bar[stage], (i / p.stages) & 1u);   // TMA bytes of this slice have landed

            const unsigned char *st = ring + (size_t)stage * L.stage_bytes;
            const uint32_t *rp = reinterpret_cast<const uint32_t *>(st + L.rp_off);
            const uint32_t s_all = rp[0], e_all = rp[nr];
            const uint32_t base_s = s_all & ~3u;
            const bool fast = ((e_all - base_s + 3u) & ~3u) <= p.cap;
            // entry k of the matrix lives at ci[k - base] / va[k - base]
            const uint32_t *__restrict__ ci = fast ? reinterpret_cast<const uint32_t *>(st + L.idx_off) : p.col_idx;
            const T *__restrict__ va = fast ? reinterpret_cast<const T *>(st + L.vals_off) : vals;
            const uint32_t base = fast ? base_s : 0u;

            if constexpr (G == 32) {
                // ======== one flat entry stream per warp ========
                Lane<T, V> acc[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) acc[t].zero();                   // T::default()  sparse.rs:434
                uint32_t rr = 0;                                               // row being accumulated (slice-local)
                uint32_t row_end = rp[1];
                uint32_t closed = 0;
                auto close_row = [&]() {
                    T *crow = c_lane + (size_t)(row0 + rr) * p.ldc;
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        if (FULLN || col_ok[t]) acc[t].store(crow + t * G * V, streaming);
                        acc[t].zero();
                    }
                    ++rr;
                    row_end = rp[min(rr + 1u, nr)];
                    if (sync_rows && ++closed == sync_rows) {
                        closed = 0;
                        --barriers_left;
                        cta_bar_sync(cta_threads);
                    }
                };
                uint32_t k = s_all;
                for (; k + U <= e_all; k += U) {
                    Lane<T, V> b[U][NT];
#pragma unroll
                    for (int u = 0; u < U; ++u)   // all gathers of the chunk first (memory-level parallelism) ...
                        load_brow<T, V, NT, FULLN>(b[u], b_lane + (size_t)ci[k + u - base] * p.ldb, col_ok, G);
#pragma unroll
                    for (int u = 0; u < U; ++u) {   // ... then consume them strictly in stored order
                        while (k + u == row_end && rr + 1 < nr) close_row();
                        fma_row<T, V, NT>(acc, b[u], va[k + u - base]);
                    }
                }
                if (k < e_all) {
                    const uint32_t rem = e_all - k;
                    Lane<T, V> b[U][NT];
#pragma unroll
                    for (int u = 0; u < U - 1; ++u)
                        if ((uint32_t)u < rem) load_brow<T, V, NT, FULLN>(b[u], b_lane + (size_t)ci[k + u - base] * p.ldb, col_ok, G);
#pragma unroll
                    for (int u = 0; u < U - 1; ++u)
                        if ((uint32_t)u < rem) {
                            while (k + u == row_end && rr + 1 < nr) close_row();
                            fma_row<T, V, NT>(acc, b[u], va[k + u - base]);
                        }
                }
                while (rr < nr) close_row();   // the last row with entries, then trailing empty rows
            } else {
                // ======== 32/G rows side by side, row by row ========
                uint32_t closed = 0;
                for (uint32_t r = grp; r < p.R; r += RPP) {   // uniform trip count: barriers stay aligned
                    if (r < nr) {
                        const uint32_t s = rp[r], e = rp[r + 1];
                        Lane<T, V> acc[NT];
#pragma unroll
                        for (int t = 0; t < NT; ++t) acc[t].zero();
                        uint32_t k = s;
                        for (; k + U <= e; k += U) {
                            Lane<T, V> b[U][NT];
#pragma unroll
                            for (int u = 0; u < U; ++u)
                                load_brow<T, V, NT, FULLN>(b[u], b_lane + (size_t)ci[k + u - base] * p.ldb, col_ok, G);
#pragma unroll
                            for (int u = 0; u < U; ++u) fma_row<T, V, NT>(acc, b[u], va[k + u - base]);
                        }
                        if (k < e) {
                            const uint32_t rem = e - k;
                            Lane<T, V> b[U][NT];
#pragma unroll
                            for (int u = 0; u < U - 1; ++u)
                                if ((uint32_t)u < rem) load_brow<T, V, NT, FULLN>(b[u], b_lane + (size_t)ci[k + u - base] * p.ldb, col_ok, G);
#pragma unroll
                            for (int u = 0; u < U - 1; ++u)
                                if ((uint32_t)u < rem) fma_row<T, V, NT>(acc, b[u], va[k + u - base]);
                        }
                        T *crow = c_lane + (size_t)(row0 + r) * p.ldc;
#pragma unroll
                        for (int t = 0; t < NT; ++t)
                            if (FULLN || col_ok[t]) acc[t].store(crow + t * G * V, streaming);
                    }
                    if (sync_rows) {
                        closed += RPP;
                        if (closed >= sync_rows) {
                            closed = 0;
                            --barriers_left;
                            __syncwarp();
                            cta_bar_sync(cta_threads);
                        }
                    }
                }
            }
        }
        // slices past the end of the matrix (and short last slices) still meet the other warps
        for (; barriers_left; --barriers_left) {
            __syncwarp();
            cta_bar_sync(cta_threads);
        }
    }
}

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
template <typename T, int V, int G, int NT> static const void *row_kernel_ptr(bool fulln)
{
    return fulln ? reinterpret_cast<const void *>(&spmm_rows_kernel<T, V, G, NT, true>)
                 : reinterpret_cast<const void *>(&spmm_rows_kernel<T, V, G, NT, false>);
}

template <typename T, int V> static const void *row_kernel_select_gnt(int G, int NT, bool fulln)
{
    if (G == 32) {
        switch (NT) {
            case 1: return row_kernel_ptr<T, V, 32, 1>(fulln);
            case 2: return row_kernel_ptr<T, V, 32, 2>(fulln);
            case 4: return row_kernel_ptr<T, V, 32, 4>(fulln);
        }
        return nullptr;
    }
    if (NT != 1) return nullptr;
    switch (G) {
        case 16: return row_kernel_ptr<T, V, 16, 1>(fulln);
        case 8: return row_kernel_ptr<T, V, 8, 1>(fulln);
        case 4: return row_kernel_ptr<T, V, 4, 1>(fulln);
        case 2: return row_kernel_ptr<T, V, 2, 1>(fulln);
        case 1: return row_kernel_ptr<T, V, 1, 1>(fulln);
    }
    return nullptr;
}

static const void *row_kernel_select(int dtype, Shape sh, uint32_t n)
{
    const bool fulln = n == (uint32_t)(sh.V * sh.G * sh.NT);
    if (dtype == BSM_F64) {
        if (sh.V == 1) return row_kernel_select_gnt<double, 1>(sh.G, sh.NT, fulln);
        if (sh.V == 2) return row_kernel_select_gnt<double, 2>(sh.G, sh.NT, fulln);
    } else {
        if (sh.V == 1) return row_kernel_select_gnt<float, 1>(sh.G, sh.NT, fulln);
        if (sh.V == 2) return row_kernel_select_gnt<float, 2>(sh.G, sh.NT, fulln);
        if (sh.V == 4) return row_kernel_select_gnt<float, 4>(sh.G, sh.NT, fulln);
    }
    return nullptr;
}

int row_kernel_occupancy(int dtype, Shape sh, uint32_t n, int block, size_t smem, int *blocks_per_sm)
{
    const void *k = row_kernel_select(dtype, sh, n);
    if (!k) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_rows: no kernel for this lane shape");
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, block, smem));
    return BSM_OK;
}

int launch_spmm_rows(int dtype, Shape sh, const RowParams &p, int grid, int block, size_t smem, cudaStream_t stream)
{
    const void *k = row_kernel_select(dtype, sh, p.n);
    if (!k) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_rows: no kernel for this lane shape");
    if (p.stages < 1 || p.stages > (uint32_t)kMaxStages) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_rows: stages out of range");
    if (p.R == 0 || p.R % 4 || p.P % p.R || (p.sync_rows && p.R % p.sync_rows))
        return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_rows: inconsistent slice geometry");
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RowParams pc = p;
    void *args[] = {&pc};
    BSM_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(block), args, smem, stream));
    count_launch();
    return BSM_OK;
}

}  // namespace bsm
