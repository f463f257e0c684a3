// line_length.h — the stencil line length one row suggests; shared by the device statistics kernel (convert.cu) and
// the host export bsm_line_length_of_row (CPU-tested in tests/test_partition_and_gen.py).
//
// Stencil-like matrices (Laplacians on a grid) touch B at fixed column offsets from the diagonal: +-1, +-nx, +-nx*ny.
// The smallest offset above 1 ("row stride") tells the vector kernel how many consecutive rows one warp should own so
// that the warps of a CTA sweep adjacent grid lines and share those B rows through L1. A performance hint only: any
// value is correct.
// Line length seen from one row: the smallest distance > 1 of a stored column from the diagonal (NOT from the row's
// median column: on a grid boundary row the median is a neighbour and the result is off by one — 4095 for a line of
// 4096). A box stencil (9- / 27-point) also stores the neighbours of its line neighbour — distances nx-1, nx, nx+1 —
// and yields nx. 0 when the row has no such column.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BSM_HD __host__ __device__
#else
#define BSM_HD
#endif

namespace bsm {

BSM_HD inline uint32_t line_length_of_row(const uint32_t *cols, uint32_t len, uint64_t diag)
{
    uint64_t stride = 0;
    for (uint32_t i = 0; i < len; ++i) {
        const uint64_t d = cols[i] > diag ? cols[i] - diag : diag - cols[i];
        if (d > 1 && (stride == 0 || d < stride)) stride = d;
    }
    if (stride == 0 || stride > 0xFFFFFFF0ull) return 0;
    bool plus1 = false, plus2 = false;
    for (uint32_t i = 0; i < len; ++i) {
        const uint64_t d = cols[i] > diag ? cols[i] - diag : diag - cols[i];
        plus1 |= d == stride + 1;
        plus2 |= d == stride + 2;
    }
    return (uint32_t)(plus1 && plus2 ? stride + 1 : stride);
}

// layout of the statistics scratch (u32 words) filled by csr_stats_kernel and read back ONCE per matrix
constexpr uint32_t kStatMaxLen = 0, kStatBadRowPtr = 1, kStatColMin = 2, kStatColMax = 3, kStatNarrowColBad = 4,
                   kStatNarrowRowBad = 5, kStatVoters = 6, kStatHist = 8;
constexpr uint32_t kStrideMin = 16, kStrideMax = 16384;          // line lengths the vector kernel can use
constexpr uint32_t kStatWords = kStatHist + kStrideMax + 1;      // histogram bin s = rows voting for line length s

}  // namespace bsm
