// spmm_rows_inst.cuh — instantiation table of spmm_rows_kernel for one element type (included by
// spmm_rows_f64.cu / spmm_rows_f32.cu so the two compile in parallel).
#pragma once
#include "spmm_rows_kernel.cuh"

namespace bsm {

constexpr int row_default_u(int NT) { return NT >= 4 ? 2 : (NT == 2 ? 4 : 8); }

template <typename T, int V, int G, int NT, bool FULLN, int U, int MAXT, int MINB, bool STAGED = true, bool VECA = true, bool MULTI = false>
static const void *rk()
{
    return reinterpret_cast<const void *>(&spmm_rows_kernel<T, V, G, NT, FULLN, U, MAXT, MINB, STAGED, VECA, MULTI>);
}

// Register-budget flavours (`flavour` argument of the selectors):
//   0: CTAs of up to 512 threads, 1 per SM (<= 128 registers)       — the default
//   1: same, gather window twice as deep
//   2: CTAs of up to 256 threads, 3 per SM (<= 85 registers)
//   3: (retired: 4 CTAs of 256 threads at <= 64 registers spilled and measured slower; runs as 2)
//   4: flavour 2 with scalar (one LDS per entry) instead of LDS.128 reads of col_idx / values
//   5, 6: one CTA of up to 768 threads per SM (<= 85 registers), LDS.128 / scalar reads
//   7: flavour 6 with a window of 10 gathers when a lane holds one register tile (else flavour 4)
//  -1: flavour 0 without TMA staging of col_idx / values (slices longer than a stage can hold)
// G == 32 shapes with all columns valid exist in all; everything else in flavours 0 and -1 only.
// `multi` (scatter of C rows to peer GPUs) exists for the default flavour of every shape and for the unstaged one.
template <typename T, int V, int NT> static const void *rk_wide(bool fulln, int flavour, bool multi)
{
    constexpr int U1 = row_default_u(NT), U2 = 2 * U1;
    if (multi) {
        if (flavour < 0) return fulln ? rk<T, V, 32, NT, true, U1, 512, 1, false, true, true>() : rk<T, V, 32, NT, false, U1, 512, 1, false, true, true>();
        if (!fulln) return rk<T, V, 32, NT, false, U1, 512, 1, true, true, true>();
        return flavour == 4 ? rk<T, V, 32, NT, true, U1, 256, 3, true, false, true>() : rk<T, V, 32, NT, true, U1, 256, 3, true, true, true>();
    }
    if (flavour < 0) return fulln ? rk<T, V, 32, NT, true, U1, 512, 1, false>() : rk<T, V, 32, NT, false, U1, 512, 1, false>();
    if (fulln) {
        switch (flavour) {
            case 1: return rk<T, V, 32, NT, true, U2, 512, 1>();
            case 2: return rk<T, V, 32, NT, true, U1, 256, 3>();
            case 4: return rk<T, V, 32, NT, true, U1, 256, 3, true, false>();   // flavour 2 with scalar A-stream reads
            case 5: return rk<T, V, 32, NT, true, U1, 768, 1>();               // ONE CTA of 24 warps per SM (24 adjacent lines share L1)
            case 6: return rk<T, V, 32, NT, true, U1, 768, 1, true, false>();  //   " with scalar A-stream reads
            case 7:   // one tile per lane: window of 10 gathers — as deep as 85 registers allow without spilling
                if constexpr (NT == 1) return rk<T, V, 32, NT, true, 10, 768, 1, true, false>();
                else return rk<T, V, 32, NT, true, U1, 256, 3, true, false>();   // (deeper windows measured slower with several tiles)
        }
        return rk<T, V, 32, NT, true, U1, 512, 1>();
    }
    return rk<T, V, 32, NT, false, U1, 512, 1>();
}

template <typename T, int V, int G> static const void *rk_narrow(bool fulln, int flavour, bool multi)
{
    if (multi) {
        if (flavour < 0) return fulln ? rk<T, V, G, 1, true, 8, 512, 1, false, true, true>() : rk<T, V, G, 1, false, 8, 512, 1, false, true, true>();
        return fulln ? rk<T, V, G, 1, true, 8, 512, 1, true, true, true>() : rk<T, V, G, 1, false, 8, 512, 1, true, true, true>();
    }
    if (flavour < 0) return fulln ? rk<T, V, G, 1, true, 8, 512, 1, false>() : rk<T, V, G, 1, false, 8, 512, 1, false>();
    if (flavour == 4) return fulln ? rk<T, V, G, 1, true, 8, 512, 1, true, false>() : rk<T, V, G, 1, false, 8, 512, 1, true, false>();   // scalar A reads
    return fulln ? rk<T, V, G, 1, true, 8, 512, 1>() : rk<T, V, G, 1, false, 8, 512, 1>();
}

// G < 32 with several register tiles per lane (lane groups walking their own flat streams): full-width
// shapes only, scalar A-stream reads, 3 CTAs x 8 warps (flavour 4, the default) or one CTA of 24 warps (6)
template <typename T, int V, int G, int NT> static const void *rk_grouped(bool fulln, int flavour, bool multi)
{
    constexpr int U1 = row_default_u(NT);
    if (!fulln || multi) return nullptr;
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
inline int row_flavour_max_threads(int flavour) { return flavour >= 2 ? 256 : 512; }

}  // namespace bsm
