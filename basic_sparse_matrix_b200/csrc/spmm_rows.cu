// spmm_rows.cu — vector-CSR SpMM/SpMV for short and regular rows (sm_100a).
//
// Replaces the loop nest of Csr::mul_dense, /root/reference/src/sparse.rs:431-444:
//     for row { row = get_row_compact(row); for out_col { value = 0;
//         for entry in row (stored order) { value = value + entry.v * B[entry.col][out_col] } } }
//
// Design (B200):
//   * persistent CTAs (grid = SMs x resident CTAs); CTA i takes row batches i, i+grid, ... so all
//     SMs sweep the matrix as one narrow front (keeps the stencil neighbours of B in the 126 MB L2);
//   * one producer lane per CTA streams each batch's contiguous slice of col_idx / values (and its
//     row_ptr slice) into shared memory with TMA bulk copies (cp.async.bulk -> UBLKCP) through a
//     multi-stage full/empty mbarrier ring; the A stream never occupies registers or L1;
//   * compute warps: a group of G lanes owns one output row, every lane owns V consecutive
//     columns per register tile (V*sizeof(T) = up to 16 bytes -> 128-bit coalesced B-row loads via
//     the read-only path), and walks the row's entries IN STORED ORDER with a separately rounded
//     multiply and add -> bit-identical to the reference's sequential sum for any input;
//   * C rows are written once with streaming stores (st.global.cs) so they do not evict B from L2.
//
// Batches whose entry count exceeds the stage capacity (irregular matrices forced onto this
// kernel) fall back to reading col_idx/values straight from global memory — same arithmetic.
#include "bsm_common.cuh"
#include "kernels.h"

namespace bsm {

constexpr int kMaxStages = 8;

struct RowSmemLayout {
    uint32_t vals_off, idx_off, rp_off, stage_bytes;
};
__host__ __device__ inline RowSmemLayout row_layout(uint32_t cap, uint32_t rb, uint32_t tsize)
{
    RowSmemLayout l;
    l.vals_off = 0;
    l.idx_off = cap * tsize;                  // cap % 4 == 0 -> 16-byte aligned
    l.rp_off = l.idx_off + cap * 4;
    l.stage_bytes = l.rp_off + (rb + 4) * 4;  // rb % 4 == 0
    return l;
}

size_t row_kernel_smem_bytes(int dtype, const RowParams &p)
{
    return (size_t)row_layout(p.cap, p.rb, (uint32_t)dtype_size(dtype)).stage_bytes * p.stages;
}

// Accumulate one row's entries [s,e) in stored order. `ci`/`va` are indexed by (entry - base).
template <typename T, int V, int G, int NT, int U>
__device__ __forceinline__ void accumulate_row(const uint32_t *__restrict__ ci, const T *__restrict__ va, uint32_t base,
                                               uint32_t s, uint32_t e, const T *__restrict__ b_lane, uint32_t ldb,
                                               const bool (&col_ok)[NT], uint32_t row, uint32_t far_thr,
                                               Lane<T, V> (&acc)[NT])
{
    for (uint32_t e0 = s; e0 < e; e0 += U) {
        Lane<T, V> b[U][NT];
        // issue all B-row gathers of this group of entries first (memory-level parallelism) ...
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e0 + u < e) {
                const uint32_t c = ci[e0 + u - base];
                const T *brow = b_lane + (size_t)c * ldb;
                // far columns (stencil planes) are used once per SM: keep them out of L1
                const uint32_t dist = c > row ? c - row : row - c;
                const bool na = far_thr != 0 && dist > far_thr;
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    if (col_ok[t]) b[u][t].load(brow + t * G * V, na);
            }
        }
        // ... then consume them strictly in stored order
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e0 + u < e) {
                const T a = va[e0 + u - base];
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        acc[t].x[i] = mul_add<false>(a, b[u][t].x[i], acc[t].x[i]);   // sparse.rs:438-439
            }
        }
    }
}

template <typename T, int V, int G, int NT>
__global__ void __launch_bounds__(288, 2) spmm_rows_kernel(const RowParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ uint32_t meta_base[kMaxStages];
    __shared__ uint32_t meta_fast[kMaxStages];

    constexpr int U = NT >= 4 ? 2 : (NT == 2 ? 4 : 8);
    constexpr int ROWS_PER_PASS = 32 / G;

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t num_compute_warps = (blockDim.x >> 5) - 1;
    const RowSmemLayout L = row_layout(p.cap, p.rb, sizeof(T));

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], num_compute_warps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const T *__restrict__ vals = static_cast<const T *>(p.vals);

    if (warp == num_compute_warps) {
        // ================= producer: one lane issues the TMA bulk copies =================
        if (lane == 0) {
            const uint64_t policy = (p.flags & BSM_TUNE_A_EVICT_FIRST) ? l2_policy_evict_first() : l2_policy_evict_normal();
            uint32_t stage = 0, phase = 0;
            uint32_t b = blockIdx.x;
            uint32_t s_cur = 0, e_cur = 0;
            if (b < p.num_batches) {
                const uint32_t r0 = b * p.rb;
                const uint32_t r1 = min(r0 + p.rb, p.rows);
                s_cur = __ldg(p.row_ptr + r0);
                e_cur = __ldg(p.row_ptr + r1);
            }
            for (; b < p.num_batches; b += gridDim.x) {
                const uint32_t r0 = b * p.rb;
                const uint32_t r1 = min(r0 + p.rb, p.rows);
                // prefetch the next batch's entry range while this one is being issued
                uint32_t s_nxt = 0, e_nxt = 0;
                const uint32_t bn = b + gridDim.x;
                if (bn < p.num_batches) {
                    const uint32_t n0 = bn * p.rb;
                    const uint32_t n1 = min(n0 + p.rb, p.rows);
                    s_nxt = __ldg(p.row_ptr + n0);
                    e_nxt = __ldg(p.row_ptr + n1);
                }
                mbar_wait(&empty_bar[stage], phase ^ 1);   // slot free (passes at once on the first lap)

                unsigned char *st = smem + (size_t)stage * L.stage_bytes;
                const uint32_t base = s_cur & ~3u;                  // 16-byte aligned start for u32 and T
                const uint32_t cnt = (e_cur - base + 3u) & ~3u;     // entries, multiple of 4
                const uint32_t cnt_r = (r1 - r0 + 1u + 3u) & ~3u;   // row_ptr slice rp[r0..r1]
                const bool fast = cnt <= p.cap;
                meta_base[stage] = base;
                meta_fast[stage] = fast ? 1u : 0u;
                const uint32_t bytes = cnt_r * 4u + ((fast && cnt) ? cnt * (4u + (uint32_t)sizeof(T)) : 0u);
                mbar_arrive_expect_tx(&full_bar[stage], bytes);
                bulk_g2s(st + L.rp_off, p.row_ptr + r0, cnt_r * 4u, &full_bar[stage], policy);
                if (fast && cnt) {
                    bulk_g2s(st + L.idx_off, p.col_idx + base, cnt * 4u, &full_bar[stage], policy);
                    bulk_g2s(st + L.vals_off, vals + base, cnt * (uint32_t)sizeof(T), &full_bar[stage], policy);
                }
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
                s_cur = s_nxt;
                e_cur = e_nxt;
            }
        }
        return;
    }

    // ================= compute warps =================
    const uint32_t grp = lane / G;    // which of the warp's concurrent rows
    const uint32_t gl = lane % G;     // lane inside the group
    bool col_ok[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) col_ok[t] = (uint32_t)((t * G + gl) * V) < p.n;
    const T *__restrict__ b_lane = static_cast<const T *>(p.B) + gl * V;
    T *__restrict__ c_lane = static_cast<T *>(p.C) + gl * V;
    const bool streaming = (p.flags & BSM_TUNE_C_STREAMING) != 0;

    uint32_t stage = 0, phase = 0;
    for (uint32_t b = blockIdx.x; b < p.num_batches; b += gridDim.x) {
        const uint32_t r0 = b * p.rb;
        const uint32_t nrows = min(p.rb, p.rows - r0);
        mbar_wait(&full_bar[stage], phase);   // TMA bytes of this stage have landed

        const unsigned char *st = smem + (size_t)stage * L.stage_bytes;
        const uint32_t *rp = reinterpret_cast<const uint32_t *>(st + L.rp_off);
        const uint32_t *idx_s = reinterpret_cast<const uint32_t *>(st + L.idx_off);
        const T *val_s = reinterpret_cast<const T *>(st + L.vals_off);
        const uint32_t base = meta_base[stage];
        const bool fast = meta_fast[stage] != 0;

        const uint32_t w_begin = warp * p.rows_per_warp;
        const uint32_t w_end = min(w_begin + p.rows_per_warp, nrows);
        for (uint32_t rr = w_begin + grp; rr < w_end; rr += ROWS_PER_PASS) {
            const uint32_t row = r0 + rr;
            const uint32_t s = rp[rr];
            const uint32_t e = rp[rr + 1];
            Lane<T, V> acc[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[t].zero();                                   // T::default()  sparse.rs:434
            if (fast)
                accumulate_row<T, V, G, NT, U>(idx_s, val_s, base, s, e, b_lane, p.ldb, col_ok, row, p.far_thr, acc);
            else
                accumulate_row<T, V, G, NT, U>(p.col_idx, vals, 0u, s, e, b_lane, p.ldb, col_ok, row, p.far_thr, acc);
            T *crow = c_lane + (size_t)row * p.ldc;
#pragma unroll
            for (int t = 0; t < NT; ++t)
                if (col_ok[t]) acc[t].store(crow + t * G * V, streaming);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);   // this warp is done reading the stage
        if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
template <typename T, int V, int G, int NT> static const void *row_kernel_ptr()
{
    return reinterpret_cast<const void *>(&spmm_rows_kernel<T, V, G, NT>);
}

template <typename T, int V> static const void *row_kernel_select_gnt(int G, int NT)
{
    if (G == 32) {
        switch (NT) {
            case 1: return row_kernel_ptr<T, V, 32, 1>();
            case 2: return row_kernel_ptr<T, V, 32, 2>();
            case 4: return row_kernel_ptr<T, V, 32, 4>();
        }
        return nullptr;
    }
    if (NT != 1) return nullptr;
    switch (G) {
        case 16: return row_kernel_ptr<T, V, 16, 1>();
        case 8: return row_kernel_ptr<T, V, 8, 1>();
        case 4: return row_kernel_ptr<T, V, 4, 1>();
        case 2: return row_kernel_ptr<T, V, 2, 1>();
        case 1: return row_kernel_ptr<T, V, 1, 1>();
    }
    return nullptr;
}

static const void *row_kernel_select(int dtype, Shape sh)
{
    if (dtype == BSM_F64) {
        if (sh.V == 1) return row_kernel_select_gnt<double, 1>(sh.G, sh.NT);
        if (sh.V == 2) return row_kernel_select_gnt<double, 2>(sh.G, sh.NT);
    } else {
        if (sh.V == 1) return row_kernel_select_gnt<float, 1>(sh.G, sh.NT);
        if (sh.V == 2) return row_kernel_select_gnt<float, 2>(sh.G, sh.NT);
        if (sh.V == 4) return row_kernel_select_gnt<float, 4>(sh.G, sh.NT);
    }
    return nullptr;
}

int row_kernel_occupancy(int dtype, Shape sh, int block, size_t smem, int *blocks_per_sm)
{
    const void *k = row_kernel_select(dtype, sh);
    if (!k) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_rows: no kernel for this lane shape");
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutDefault));
    BSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, block, smem));
    return BSM_OK;
}

int launch_spmm_rows(int dtype, Shape sh, const RowParams &p, int grid, int block, size_t smem, cudaStream_t stream)
{
    const void *k = row_kernel_select(dtype, sh);
    if (!k) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_rows: no kernel for this lane shape");
    if (p.stages < 1 || p.stages > (uint32_t)kMaxStages) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_rows: stages out of range");
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RowParams pc = p;
    void *args[] = {&pc};
    BSM_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(block), args, smem, stream));
    count_launch();
    return BSM_OK;
}

}  // namespace bsm
