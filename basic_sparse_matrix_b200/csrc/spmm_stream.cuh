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

}  // namespace bsm
