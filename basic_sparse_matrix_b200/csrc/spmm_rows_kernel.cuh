// spmm_rows_kernel.cuh — vector-CSR SpMM/SpMV for short and regular rows (sm_100a).
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
#pragma once
#include "bsm_common.cuh"
#include "kernels.h"
#include "spmm_stream.cuh"

namespace bsm {

constexpr int kMaxStages = 8;

// Which shapes walk ONE flat entry stream per lane group (the gather window never drains at a row end) instead of 32/G rows side
// by side, row by row: a full warp per row, several register tiles per lane — and, as a kernel variant of its own (FLAT), narrow
// one-tile shapes on short regular rows (64-byte output rows on a 7-point stencil: 1.26 -> 0.94 ms).
template <int G, int NT, bool FLAT> constexpr bool kFlatStream = G == 32 || NT > 1 || FLAT;

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

// One slice of one warp: rows [row0, row0+nr), entry k of the matrix at ci[k - base] / va[k - base]
// (shared memory when the slice was staged by TMA, else the global arrays with base = 0).
// VECA: col_idx / values sit in a 16-byte aligned shared-memory stage (padded past the slice), so the A
// stream is read four columns / two or four values per LDS.128 instead of one scalar LDS per entry.
// MULTI: every C row is also written to p.n_peers further destinations (the full result buffers of the
// other GPUs, mapped over NVLink): multiply and all-gather in one kernel, P2P stores instead of a
// collective. `lane_off` = byte offset of the lane's first column inside a row.
template <typename T, int V, int G, int NT, bool FULLN, int U, bool VECA, bool MULTI, bool FLAT>
__device__ __forceinline__ void process_slice(const RowParams &p, const uint32_t *__restrict__ rp, const uint32_t *__restrict__ ci,
                                              const T *__restrict__ va, uint32_t base, uint32_t row0, uint32_t nr,
                                              const char *__restrict__ b_bytes, char *__restrict__ c_bytes, const bool (&col_ok)[NT],
                                              uint32_t grp, bool streaming, uint32_t lane_off)
{
    constexpr int RPP = 32 / G;
    const uint32_t ldb_bytes = p.ldb * (uint32_t)sizeof(T);
    const uint32_t ldc_bytes = p.ldc * (uint32_t)sizeof(T);
    ci -= base;   // entry k at ci[k] / va[k]
    va -= base;
    if constexpr (kFlatStream<G, NT, FLAT>) {
        // ======== one flat entry stream per lane group ========
        // G == 32: the warp walks the whole slice. G < 32 (several register tiles per lane): the slice is cut
        // into 32/G runs of consecutive rows, one per lane group — one LDS of col_idx / values then feeds
        // 32/G rows (the A stream costs the L1 data pipe as much per entry as a quarter of a 512-byte gather).
        const uint32_t nrg = RPP == 1 ? nr : (nr + RPP - 1) / RPP;
        const uint32_t g0 = RPP == 1 ? 0u : min(grp * nrg, nr);   // this group's rows [g0, g1) of the slice
        const uint32_t g1 = RPP == 1 ? nr : min(g0 + nrg, nr);
        Lane<T, V> acc[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t].zero();                   // T::default()  sparse.rs:434
        uint32_t rr = g0;                                              // row being accumulated (slice-local)
        uint32_t row_end = rp[min(g0 + 1u, g1)];
        size_t crow = (size_t)(row0 + g0) * ldc_bytes;   // byte offset of the row inside C (and inside every peer copy)
        auto close_row = [&]() {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                if (FULLN || col_ok[t]) {
                    acc[t].store(reinterpret_cast<T *>(c_bytes + crow) + t * G * V, streaming);
                    if constexpr (MULTI)
                        for (uint32_t d = 0; d < p.n_peers; ++d)
                            acc[t].store(reinterpret_cast<T *>(p.peers[d] + lane_off + crow) + t * G * V, false);
                }
                acc[t].zero();
            }
            crow += ldc_bytes;
            ++rr;
            row_end = rp[min(rr + 1u, g1)];
        };
        stream_entries<T, V, NT, FULLN, U, VECA, false, false, false>(ci, va, rp[g0], rp[g1], b_bytes, ldb_bytes, col_ok, G, acc, [&](uint32_t k) {
            while (k == row_end) close_row();
        });
        while (rr < g1) close_row();   // the last row with entries, then trailing empty rows
    } else {
        // ======== 32/G rows side by side, row by row ========
        for (uint32_t r = grp; r < nr; r += RPP) {
            const uint32_t s = rp[r], e = rp[r + 1];
            Lane<T, V> acc[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[t].zero();
            stream_entries<T, V, NT, FULLN, U, VECA, false, true, false>(ci, va, s, e, b_bytes, ldb_bytes, col_ok, G, acc, [](uint32_t) {});
            const size_t crow = (size_t)(row0 + r) * ldc_bytes;
#pragma unroll
            for (int t = 0; t < NT; ++t)
                if (FULLN || col_ok[t]) {
                    acc[t].store(reinterpret_cast<T *>(c_bytes + crow) + t * G * V, streaming);
                    if constexpr (MULTI)
                        for (uint32_t d = 0; d < p.n_peers; ++d)
                            acc[t].store(reinterpret_cast<T *>(p.peers[d] + lane_off + crow) + t * G * V, false);
                }
        }
    }
}

// U = B-row gathers kept in flight per lane group (a rolling window: entry k+U is requested as soon
// as entry k has been consumed); MAXT / MINB = launch bounds (threads per CTA, CTAs per SM the register allocation must allow).
// STAGED = col_idx / values of every slice fit the TMA stage (the host guarantees it from the longest
// row); the unstaged variant reads them from global memory and stages only the row_ptr windows.
template <typename T, int V, int G, int NT, bool FULLN, int U, int MAXT, int MINB, bool STAGED, bool VECA = true, bool MULTI = false, bool FLAT = false>
__global__ void __launch_bounds__(MAXT, MINB) spmm_rows_kernel(const RowParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];

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
    const uint32_t spw = (p.P + p.R - 1) / p.R; // slices per warp per super-batch (the last one is short when R does not divide P)
    const uint32_t my_supers = blockIdx.x < p.num_super ? (p.num_super - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const uint32_t my_slices = my_supers * spw;

    // Two renditions of the slice loop, chosen at compile time. Wide shapes (G == 32, or several tiles per lane) have long slices and
    // a tight register budget: the slice position is recomputed from the slice index (measured 3-4 % faster
    // there than cursors). Narrow shapes (G < 32) have short slices — a few row passes — and registers to
    // spare: incremental cursors, no divisions in the loop (SpMV 0.078 -> 0.063 ms).
    // P (rows per warp) is the line length of a stencil matrix and need be a multiple of neither R nor 4: the
    // row_ptr window of a slice is copied from the 16-byte aligned index below its first row.
    if constexpr (kFlatStream<G, NT, FLAT>) {
        // first row and row count of this warp's i-th slice (0 rows: past the end of the matrix)
        auto slice_pos = [&](uint32_t i, uint64_t &r0, uint32_t &nr) {
            const uint32_t k = i / spw, t = i - k * spw;
            r0 = (uint64_t)(blockIdx.x + k * gridDim.x) * S + (uint64_t)warp * p.P + (uint64_t)t * p.R;
            const uint32_t in_p = min(p.R, p.P - t * p.R);
            nr = r0 < p.rows ? (uint32_t)min((uint64_t)in_p, p.rows - r0) : 0u;
        };

        // ---- producer side (lane 0): TMA bulk copies of one slice into ring stage i % stages --------
        const uint64_t policy = (p.flags & BSM_TUNE_A_EVICT_FIRST) ? l2_policy_evict_first() : l2_policy_evict_normal();
        uint32_t pf_s = 0, pf_e = 0;                // entry range of the next slice to issue (prefetched)
        auto prefetch_bounds = [&](uint32_t i) {
            if (i < my_slices) {
                uint64_t r0;
                uint32_t nr;
                slice_pos(i, r0, nr);
                if (nr) {
                    pf_s = __ldg(p.row_ptr + r0);
                    pf_e = __ldg(p.row_ptr + r0 + nr);
                }
            }
        };
        auto issue = [&](uint32_t i) {
            // only lane 0 calls this
            uint64_t r0;
            uint32_t nr;
            slice_pos(i, r0, nr);
            if (nr) {
                const uint32_t stage = i % p.stages;
                unsigned char *st = ring + (size_t)stage * L.stage_bytes;
                const uint32_t base = pf_s & ~3u;                     // 16-byte aligned start for u32 and T
                const uint32_t cnt = STAGED ? ((pf_e - base + 3u) & ~3u) : 0u;   // entries, multiple of 4
                if (STAGED && cnt > p.cap) __trap();   // the host sizes cap from the longest row; never overrun the stage
                const uint32_t a0 = (uint32_t)(r0 & 3u);              // row_ptr window from the aligned index below r0
                const uint32_t cnt_r = (a0 + nr + 1u + 3u) & ~3u;     // rp[r0 - a0 .. r0 + nr], at most R + 4 entries
                mbar_arrive_expect_tx(&full_bar[stage], cnt_r * 4u + cnt * (4u + (uint32_t)sizeof(T)));
                bulk_g2s(st + L.rp_off, p.row_ptr + (r0 - a0), cnt_r * 4u, &full_bar[stage], policy);
                if (cnt) {
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
        const char *__restrict__ b_bytes = reinterpret_cast<const char *>(static_cast<const T *>(p.B) + gl * V);
        char *__restrict__ c_bytes = reinterpret_cast<char *>(static_cast<T *>(p.C) + gl * V);
        const bool streaming = (p.flags & BSM_TUNE_C_STREAMING) != 0;

        for (uint32_t i = 0; i < my_slices; ++i) {
            __syncwarp();   // every lane is done reading the stage that is refilled next
            if (lane == 0 && i + p.stages - 1 < my_slices) issue(i + p.stages - 1);

            uint64_t row0_64;
            uint32_t nr;
            slice_pos(i, row0_64, nr);
            if (nr) {
                const uint32_t row0 = (uint32_t)row0_64;
                const uint32_t stage = i % p.stages;
                mbar_wait(&full_bar[stage], (i / p.stages) & 1u);   // TMA bytes of this slice have landed

                const unsigned char *st = ring + (size_t)stage * L.stage_bytes;
                const uint32_t *rp = reinterpret_cast<const uint32_t *>(st + L.rp_off) + (row0 & 3u);
                if constexpr (STAGED)
                    process_slice<T, V, G, NT, FULLN, U, VECA, MULTI, FLAT>(p, rp, reinterpret_cast<const uint32_t *>(st + L.idx_off),
                                                                reinterpret_cast<const T *>(st + L.vals_off), rp[0] & ~3u, row0, nr, b_bytes,
                                                                c_bytes, col_ok, grp, streaming, gl * V * (uint32_t)sizeof(T));
                else
                    process_slice<T, V, G, NT, FULLN, U, false, MULTI, FLAT>(p, rp, p.col_idx, vals, 0u, row0, nr, b_bytes, c_bytes, col_ok, grp,
                                                                       streaming, gl * V * (uint32_t)sizeof(T));
            }
        }
    } else {
        // Slice bookkeeping. Narrow shapes (G < 32) have short slices (a few row passes) and registers to
        // spare: incremental cursors, no divisions in the loop. Wide shapes have long slices and a tight
        // register budget: the slice position is recomputed from the slice index.
        constexpr bool kCursor = true;
        const uint64_t sb_step = (uint64_t)gridDim.x * S;
        const uint64_t warp_row = (uint64_t)blockIdx.x * S + (uint64_t)warp * p.P;
        auto slice_row0 = [&](uint32_t i) -> uint64_t {   // first row of this warp's i-th slice
            const uint32_t k = i / spw, t = i - k * spw;
            return warp_row + (uint64_t)k * sb_step + (uint64_t)t * p.R;
        };

        // ---- producer side (lane 0): TMA bulk copies of one slice into its ring stage -----------------
        const uint64_t policy = (p.flags & BSM_TUNE_A_EVICT_FIRST) ? l2_policy_evict_first() : l2_policy_evict_normal();
        uint64_t p_base = warp_row;                 // cursor of the next slice to issue (kCursor)
        uint32_t p_t = 0, p_stage = 0;
        uint32_t p_next = 0;                        // index of the next slice to issue
        uint32_t pf_s = 0, pf_e = 0;                // its entry range (prefetched)
        auto next_row0 = [&]() -> uint64_t { return kCursor ? p_base + (uint64_t)p_t * p.R : slice_row0(p_next); };
        auto prefetch_bounds = [&]() {
            if (p_next < my_slices) {
                const uint64_t r0 = next_row0();
                if (r0 < p.rows) {
                    const uint32_t r1 = (uint32_t)min(r0 + p.R, (uint64_t)p.rows);
                    pf_s = __ldg(p.row_ptr + r0);
                    pf_e = __ldg(p.row_ptr + r1);
                }
            }
        };
        auto issue = [&]() {
            // only lane 0 calls this, with p_next < my_slices
            const uint64_t r0 = next_row0();
            const uint32_t stage = kCursor ? p_stage : p_next % p.stages;
            if (r0 < p.rows) {
                const uint32_t nr = (uint32_t)min((uint64_t)p.R, p.rows - r0);
                unsigned char *st = ring + (size_t)stage * L.stage_bytes;
                const uint32_t base = pf_s & ~3u;                     // 16-byte aligned start for u32 and T
                const uint32_t cnt = STAGED ? ((pf_e - base + 3u) & ~3u) : 0u;   // entries, multiple of 4
                if (STAGED && cnt > p.cap) __trap();   // the host sizes cap from the longest row; never overrun the stage
                const uint32_t cnt_r = (nr + 1u + 3u) & ~3u;          // row_ptr window rp[r0 .. r0+nr]
                mbar_arrive_expect_tx(&full_bar[stage], cnt_r * 4u + cnt * (4u + (uint32_t)sizeof(T)));
                bulk_g2s(st + L.rp_off, p.row_ptr + r0, cnt_r * 4u, &full_bar[stage], policy);
                if (cnt) {
                    bulk_g2s(st + L.idx_off, p.col_idx + base, cnt * 4u, &full_bar[stage], policy);
                    bulk_g2s(st + L.vals_off, vals + base, cnt * (uint32_t)sizeof(T), &full_bar[stage], policy);
                }
            }
            ++p_next;
            if constexpr (kCursor) {
                if (++p_t == spw) {
                    p_t = 0;
                    p_base += sb_step;
                }
                if (++p_stage == p.stages) p_stage = 0;
            }
            prefetch_bounds();
        };

        if (lane == 0) {
            prefetch_bounds();
            for (uint32_t i = 0; i + 1 < p.stages && p_next < my_slices; ++i) issue();
        }

        // ---- consumer side -------------------------------------------------------------------------
        const uint32_t grp = lane / G;    // which of the warp's concurrent rows (G < 32)
        const uint32_t gl = lane % G;     // lane inside the group
        bool col_ok[NT];
    #pragma unroll
        for (int t = 0; t < NT; ++t) col_ok[t] = FULLN || (uint32_t)((t * G + gl) * V) < p.n;
        const char *__restrict__ b_bytes = reinterpret_cast<const char *>(static_cast<const T *>(p.B) + gl * V);
        char *__restrict__ c_bytes = reinterpret_cast<char *>(static_cast<T *>(p.C) + gl * V);
        const bool streaming = (p.flags & BSM_TUNE_C_STREAMING) != 0;

        uint64_t c_base = warp_row;                 // consumer cursor (kCursor)
        uint32_t c_t = 0, c_stage = 0, c_phase = 0;
        for (uint32_t i = 0; i < my_slices; ++i) {
            __syncwarp();   // every lane is done reading the stage that is refilled next
            if (lane == 0 && p_next < my_slices) issue();

            uint64_t row0_64;
            uint32_t stage, phase;
            if constexpr (kCursor) {
                row0_64 = c_base + (uint64_t)c_t * p.R;
                stage = c_stage;
                phase = c_phase;
                if (++c_t == spw) {
                    c_t = 0;
                    c_base += sb_step;
                }
                if (++c_stage == p.stages) {
                    c_stage = 0;
                    c_phase ^= 1u;
                }
            } else {
                row0_64 = slice_row0(i);
                stage = i % p.stages;
                phase = (i / p.stages) & 1u;
            }
            if (row0_64 < p.rows) {
                const uint32_t row0 = (uint32_t)row0_64;
                const uint32_t nr = min(p.R, p.rows - row0);
                mbar_wait(&full_bar[stage], phase);   // TMA bytes of this slice have landed

                const unsigned char *st = ring + (size_t)stage * L.stage_bytes;
                const uint32_t *rp = reinterpret_cast<const uint32_t *>(st + L.rp_off);
                if constexpr (STAGED)
                    process_slice<T, V, G, NT, FULLN, U, VECA, MULTI, FLAT>(p, rp, reinterpret_cast<const uint32_t *>(st + L.idx_off),
                                                                reinterpret_cast<const T *>(st + L.vals_off), rp[0] & ~3u, row0, nr, b_bytes,
                                                                c_bytes, col_ok, grp, streaming, gl * V * (uint32_t)sizeof(T));
                else
                    process_slice<T, V, G, NT, FULLN, U, false, MULTI, FLAT>(p, rp, p.col_idx, vals, 0u, row0, nr, b_bytes, c_bytes, col_ok, grp,
                                                                       streaming, gl * V * (uint32_t)sizeof(T));
            }
        }
    }
}

}  // namespace bsm
