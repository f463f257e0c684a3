// spmm_stream.cuh — the B-row gather engine shared by the vector-CSR and merge-path kernels (sm_100a).
#pragma once
#include <type_traits>

#include "bsm_common.cuh"

namespace bsm {

// B row `c` of this lane: b_bytes already points at the lane's first column
// BHINT: the gathers carry the L2 policy `bpol` (evict_last keeps hot B rows of a power-law matrix in L2;
// measured +3 % on R-MAT, -20 % on a stencil whose B window exceeds L2 — so only the merge-path kernel uses it)
template <typename T, int V, int NT, bool FULLN, bool BHINT = false>
__device__ __forceinline__ void load_brow(Lane<T, V> (&b)[NT], const char *__restrict__ b_bytes, uint32_t ldb_bytes, uint32_t c,
                                          const bool (&col_ok)[NT], int G, uint64_t bpol = 0)
{
    const T *brow = reinterpret_cast<const T *>(b_bytes + (size_t)c * ldb_bytes);   // one IMAD.WIDE
#pragma unroll
    for (int t = 0; t < NT; ++t)
        if (FULLN || col_ok[t]) b[t].template load<BHINT>(brow + t * G * V, bpol);
}

// FUSED = false: value = value + (a*b) with the product and the sum rounded separately, as the
// reference does (sparse.rs:438-439); FUSED = true: one FMA
template <bool FUSED, typename T, int V, int NT>
__device__ __forceinline__ void fma_row(Lane<T, V> (&acc)[NT], const Lane<T, V> (&b)[NT], T a)
{
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[t].x[i] = mul_add<FUSED>(a, b[t].x[i], acc[t].x[i]);
}

// The gather engine shared by the vector-CSR and merge-path kernels: walk the stored entries [s, e) of
// one lane group in stored order, entry k at ci[k] / va[k], keeping a ROLLING WINDOW of U B-row
// gathers in flight — slot u holds entry k+u until it is consumed and is refilled with entry k+u+U
// at once, so memory-level parallelism never drains at row ends. `on_entry(k)` runs right before entry
// k is accumulated into `acc` (the caller closes finished rows there).
//   VECA: ci / va sit in a 16-byte aligned shared-memory stage whose element 0 is a multiple of 4
//   entries and which is padded >= 2U+4 entries past e; the A stream is then read four columns and
//   16 bytes of values per LDS.128 instead of two scalar LDS per entry. Chunks start at multiples of
//   4; the up to three entries before s are skipped like the tail.
//   SHORT: the stream is often shorter than the window (one short row per lane group): the prologue is
//   predicated instead of filling every slot.
template <typename T, int V, int NT, bool FULLN, int U, bool VECA, bool FUSED, bool SHORT, bool BHINT, typename OnEntry>
__device__ __forceinline__ void stream_entries(const uint32_t *__restrict__ ci, const T *__restrict__ va, uint32_t s, uint32_t e,
                                               const char *__restrict__ b_bytes, uint32_t ldb_bytes, const bool (&col_ok)[NT], int G,
                                               Lane<T, V> (&acc)[NT], OnEntry &&on_entry)
{
    if (s >= e) return;
    const uint64_t bpol = BHINT ? l2_policy_evict_last() : 0ull;
    Lane<T, V> b[U][NT];
    if constexpr (VECA && U % 4 == 0) {
        constexpr int VPL = 16 / (int)sizeof(T);   // values per LDS.128
        const uint32_t k0 = s & ~3u;
        const uint32_t c_first = ci[s];
#pragma unroll
        for (int q = 0; q < U / 4; ++q) {
            const uint4 c4 = *reinterpret_cast<const uint4 *>(ci + k0 + 4 * q);
            const uint32_t cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t k = k0 + 4 * q + j;
                load_brow<T, V, NT, FULLN, BHINT>(b[4 * q + j], b_bytes, ldb_bytes, (k >= s && k < e) ? cc[j] : c_first, col_ok, G, bpol);
            }
        }
        auto chunk = [&](uint32_t kk, auto pred_tag) {
            constexpr bool PRED = decltype(pred_tag)::value;
#pragma unroll
            for (int q = 0; q < U / 4; ++q) {
                const uint4 n4 = *reinterpret_cast<const uint4 *>(ci + kk + U + 4 * q);   // columns of the refills
                const uint32_t nc[4] = {n4.x, n4.y, n4.z, n4.w};
                T a4[4];
#pragma unroll
                for (int h = 0; h < 4 / VPL; ++h)
                    *reinterpret_cast<uint4 *>(a4 + h * VPL) = *reinterpret_cast<const uint4 *>(va + kk + 4 * q + h * VPL);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = 4 * q + j;
                    const uint32_t k = kk + u;
                    if (!PRED || (k >= s && k < e)) {   // consumed strictly in stored order
                        on_entry(k);
                        fma_row<FUSED, T, V, NT>(acc, b[u], a4[j]);
                    }
                    if (!PRED || (k + U >= s && k + U < e)) load_brow<T, V, NT, FULLN, BHINT>(b[u], b_bytes, ldb_bytes, nc[j], col_ok, G, bpol);
                }
            }
        };
        uint32_t kk = k0;
        chunk(kk, std::true_type{});   // first chunk: may start before s
        kk += U;
        for (; kk + 2 * U <= e; kk += U) chunk(kk, std::false_type{});   // steady state: no bounds checks
        for (; kk < e; kk += U) chunk(kk, std::true_type{});             // drain
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if constexpr (SHORT) {   // streams often shorter than the window (SpMV rows): no redundant gathers
                if (s + u < e) load_brow<T, V, NT, FULLN, BHINT>(b[u], b_bytes, ldb_bytes, ci[s + u], col_ok, G, bpol);
            } else {                 // long streams: fill every slot unconditionally (past the end: the last entry again)
                load_brow<T, V, NT, FULLN, BHINT>(b[u], b_bytes, ldb_bytes, ci[min(s + u, e - 1u)], col_ok, G, bpol);
            }
        }
        uint32_t k = s;
        for (; k + 2 * U <= e; k += U) {   // steady state: no bounds checks
#pragma unroll
            for (int u = 0; u < U; ++u) {
                on_entry(k + u);
                fma_row<FUSED, T, V, NT>(acc, b[u], va[k + u]);
                load_brow<T, V, NT, FULLN, BHINT>(b[u], b_bytes, ldb_bytes, ci[k + u + U], col_ok, G, bpol);
            }
        }
        for (; k < e; k += U) {            // drain
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (k + u < e) {
                    on_entry(k + u);
                    fma_row<FUSED, T, V, NT>(acc, b[u], va[k + u]);
                    if (k + u + U < e) load_brow<T, V, NT, FULLN, BHINT>(b[u], b_bytes, ldb_bytes, ci[k + u + U], col_ok, G, bpol);
                }
            }
        }
    }
}

// The same engine with NEIGHBOUR-ROW REUSE for one flat stream over consecutive rows (G == 32): the B rows
// of columns `row` and `row + 1` are kept in two register sets after their first use, so the entries at
// columns row-1 / row of the NEXT rows (the tridiagonal part of a stencil or FEM matrix) are served from
// registers instead of being gathered again — 7 -> 5 gathers per row of a 3-D 7-point Laplacian. The L1
// data pipe (not HBM) is the co-limit of that workload, and a gather costs it 4 wavefronts per 512 bytes.
//   A small state machine runs at ISSUE time, U entries ahead of consumption: it tracks the two columns
//   the keep registers WILL hold when the entry is consumed, skips the gather on a match, and records per
//   window slot a 2-bit code (0 gather, 1 / 2 keep set a / b, 3 gather + keep) that the consumer obeys.
//   Both run over the same entry sequence, so they agree for any matrix; entries are still accumulated
//   strictly in stored order with the unfused multiply-add -> bit-identical to stream_entries.
// rp[0..nr] = row_ptr window of the slice, row0 = global index of its first row; on_entry as above.
template <typename T, int V, int NT, int U, typename OnEntry>
__device__ __forceinline__ void stream_entries_reuse(const uint32_t *__restrict__ ci, const T *__restrict__ va,
                                                     const uint32_t *__restrict__ rp, uint32_t nr, uint32_t row0,
                                                     const char *__restrict__ b_bytes, uint32_t ldb_bytes, int G, Lane<T, V> (&acc)[NT],
                                                     OnEntry &&on_entry)
{
    static_assert(U <= 16, "2 bits per window slot in one register");
    const uint32_t s = rp[0], e = rp[nr];
    if (s >= e) return;
    bool col_ok[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) col_ok[t] = true;
    Lane<T, V> b[U][NT], keep_a[NT], keep_b[NT];
    uint32_t codes = 0;
    uint32_t col_a = 0xFFFFFFFFu, col_b = 0xFFFFFFFFu;   // no column has this index (indices < 2^32 - 16)
    uint32_t i_rr = 0, i_row = row0, i_end = rp[1];     // row of the entry being issued
    auto issue = [&](uint32_t k, int u) {
        while (k == i_end) {
            ++i_rr;
            ++i_row;
            i_end = rp[min(i_rr + 1u, nr)];
        }
        const uint32_t c = ci[k];
        uint32_t code;
        if (c == col_a) {
            code = 1u;
        } else if (c == col_b) {
            code = 2u;
        } else {
            load_brow<T, V, NT, true, false>(b[u], b_bytes, ldb_bytes, c, col_ok, G);
            code = 0u;
            if (c - i_row <= 1u) {   // column == row or row + 1: the next rows will ask for it again
                code = 3u;
                col_b = col_a;
                col_a = c;
            }
        }
        codes = (codes & ~(3u << (2 * u))) | (code << (2 * u));
    };
    auto consume = [&](uint32_t k, int u) {
        on_entry(k);
        const T a = va[k];
        const uint32_t code = (codes >> (2 * u)) & 3u;
        if (code == 1u) {
            fma_row<false, T, V, NT>(acc, keep_a, a);
        } else if (code == 2u) {
            fma_row<false, T, V, NT>(acc, keep_b, a);
        } else {
            fma_row<false, T, V, NT>(acc, b[u], a);
            if (code == 3u) {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    keep_b[t] = keep_a[t];
                    keep_a[t] = b[u][t];
                }
            }
        }
    };
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (s + u < e) issue(s + u, u);
    uint32_t k = s;
    for (; k + 2 * U <= e; k += U) {   // steady state: no bounds checks
#pragma unroll
        for (int u = 0; u < U; ++u) {
            consume(k + u, u);
            issue(k + u + U, u);
        }
    }
    for (; k < e; k += U) {            // drain
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (k + u < e) {
                consume(k + u, u);
                if (k + u + U < e) issue(k + u + U, u);
            }
        }
    }
}

}  // namespace bsm
