// spmm_rows_inst.cuh — instantiation table of spmm_rows_kernel for one element type (included by
// spmm_rows_f64.cu / spmm_rows_f32.cu so the two compile in parallel).
#pragma once
#include "spmm_rows_kernel.cuh"

namespace bsm {

constexpr int row_default_u(int NT) { return NT >= 4 ? 2 : (NT == 2 ? 4 : 8); }

template <typename T, int V, int G, int NT, bool FULLN, int U, int MAXT, int MINB, bool STAGED = true, bool VECA = true, bool MULTI = false, bool FLAT = false>
static const void *rk()
{
    return reinterpret_cast<const void *>(&spmm_rows_kernel<T, V, G, NT, FULLN, U, MAXT, MINB, STAGED, VECA, MULTI, FLAT>);
}

// Register-budget flavours (`flavour` argument of the selectors = bsm_tuning.reg_flavour - 1):
//   0: CTAs of up to 512 threads, 1 per SM (<= 128 registers), LDS.128 reads of the staged col_idx / values
//   4: CTAs of up to 256 threads, 3 per SM (<= 85 registers), scalar reads (one LDS per entry)  — default for several tiles per lane
//   6: one CTA of up to 768 threads per SM (<= 85 registers), scalar reads                       — default for grouped lanes
//   7: flavour 6 with a window of 10 gathers when a lane holds one register tile (else flavour 4) — default for one tile per lane
//   (narrow one-tile shapes: 0, and for a row per lane 4 = scalar reads)
//   8: narrow one-tile shapes of 4 x 128-bit lanes as flat entry streams per lane group (flavour 0 otherwise) — default for 64-byte
//      output rows on short regular rows
//  -1: flavour 0 without TMA staging of col_idx / values (slices longer than a stage can hold)
// Flavours 1, 2, 3, 5 of round 1 (window twice as deep; LDS.128 reads at 3 CTAs / at 768 threads; 4 CTAs at 64 registers) lost
// every sweep they were in (profiles/r1_sweep{i,s}_*.jsonl) and are no longer built: 1 runs as 0, 2 and 3 as 4, 5 as 6.
// G == 32 shapes with all columns valid exist in all; everything else in flavours 0 and -1 only.
// `multi` (scatter of C rows to peer GPUs) exists for the default flavour of every shape and for the unstaged one.
inline int row_flavour_built(int flavour)
{
    switch (flavour) {
        case 1: return 0;
        case 2:
        case 3: return 4;
        case 5: return 6;
    }
    return flavour;
}

template <typename T, int V, int NT> static const void *rk_wide(bool fulln, int flavour, bool multi)
{
    constexpr int U1 = row_default_u(NT);
    flavour = row_flavour_built(flavour);
    // scatter variant: bound by the P2P stores over NVLink, not by the SM — one (predicated, staged) kernel per shape
    if (multi) return flavour < 0 ? nullptr : rk<T, V, 32, NT, false, U1, 512, 1, true, true, true>();
    if (flavour < 0) return rk<T, V, 32, NT, false, U1, 512, 1, false>();   // unstaged: one (predicated) variant serves full-width shapes too
    if constexpr (V * sizeof(T) == 16) {   // the register-budget flavours are built for 128-bit lanes (one-element lanes: the default)
        if (fulln) {
            switch (flavour) {
                case 4: return rk<T, V, 32, NT, true, U1, 256, 3, true, false>();
                case 6: return rk<T, V, 32, NT, true, U1, 768, 1, true, false>();  // ONE CTA of 24 warps per SM (24 adjacent lines share L1)
                case 7:   // one tile per lane: window of 10 gathers — as deep as 85 registers allow without spilling
                    if constexpr (NT == 1) return rk<T, V, 32, NT, true, 10, 768, 1, true, false>();
                    else return rk<T, V, 32, NT, true, U1, 256, 3, true, false>();   // (deeper windows measured slower with several tiles)
            }
        }
    }
    if (fulln) return rk<T, V, 32, NT, true, U1, 512, 1>();
    return rk<T, V, 32, NT, false, U1, 512, 1>();
}

template <typename T, int V, int G> static const void *rk_narrow(bool fulln, int flavour, bool multi)
{
    if (multi) return flavour < 0 ? nullptr : rk<T, V, G, 1, false, 8, 512, 1, true, true, true>();   // scatter variant (see rk_wide)
    if (flavour < 0) return rk<T, V, G, 1, false, 8, 512, 1, false>();   // unstaged: one (predicated) variant
    if constexpr (G == 1) {   // a row per lane (SpMV): scalar reads of the staged col_idx / values
        if (flavour == 4) return fulln ? rk<T, V, G, 1, true, 8, 512, 1, true, false>() : rk<T, V, G, 1, false, 8, 512, 1, true, false>();
    }
    if constexpr (G == 4 && V * sizeof(T) == 16) {   // flat entry streams (8 rows side by side never drain their windows at a row end)
        if (flavour == 8) return fulln ? rk<T, V, G, 1, true, 8, 512, 1, true, true, false, true>() : rk<T, V, G, 1, false, 8, 512, 1, true, true, false, true>();
    }
    return fulln ? rk<T, V, G, 1, true, 8, 512, 1>() : rk<T, V, G, 1, false, 8, 512, 1>();
}

// G < 32 with several register tiles per lane (lane groups walking their own flat streams): full-width
// shapes only, scalar A-stream reads, 3 CTAs x 8 warps (flavour 4, the default) or one CTA of 24 warps (6)
template <typename T, int V, int G, int NT> static const void *rk_grouped(bool fulln, int flavour, bool multi)
{
    constexpr int U1 = row_default_u(NT);
    if (!fulln || multi) return nullptr;
    flavour = row_flavour_built(flavour);
    if (flavour == 6) return rk<T, V, G, NT, true, U1, 768, 1, true, false>();
    return rk<T, V, G, NT, true, U1, 256, 3, true, false>();
}

template <typename T, int V> static const void *row_kernel_select_v(Shape sh, bool fulln, int flavour, bool multi)
{
    if (sh.G == 32) {
        switch (sh.NT) {
            case 1: return rk_wide<T, V, 1>(fulln, flavour, multi);
            case 2: return rk_wide<T, V, 2>(fulln, flavour, multi);
            case 4: return rk_wide<T, V, 4>(fulln, flavour, multi);
        }
        return nullptr;
    }
    if (sh.NT != 1) {
        if constexpr (V * sizeof(T) == 16) {   // grouped flat streams: 128-bit lanes only
            if (sh.NT == 2) {
                switch (sh.G) {
                    case 16: return rk_grouped<T, V, 16, 2>(fulln, flavour, multi);
                    case 8: return rk_grouped<T, V, 8, 2>(fulln, flavour, multi);
                    case 4: return rk_grouped<T, V, 4, 2>(fulln, flavour, multi);
                }
            } else if (sh.NT == 4) {
                switch (sh.G) {
                    case 16: return rk_grouped<T, V, 16, 4>(fulln, flavour, multi);
                    case 8: return rk_grouped<T, V, 8, 4>(fulln, flavour, multi);
                    case 4: return rk_grouped<T, V, 4, 4>(fulln, flavour, multi);
                }
            }
        }
        return nullptr;
    }
    switch (sh.G) {
        case 16: return rk_narrow<T, V, 16>(fulln, flavour, multi);
        case 8: return rk_narrow<T, V, 8>(fulln, flavour, multi);
        case 4: return rk_narrow<T, V, 4>(fulln, flavour, multi);
        case 2: return rk_narrow<T, V, 2>(fulln, flavour, multi);
        case 1: return rk_narrow<T, V, 1>(fulln, flavour, multi);
    }
    return nullptr;
}

// largest CTA (threads) a flavour was compiled for
inline int row_flavour_max_threads(int flavour) { return flavour >= 2 && flavour != 8 ? 256 : 512; }

}  // namespace bsm
