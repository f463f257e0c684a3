// planner.cu — pure host arithmetic (no CUDA calls): lane shapes, the launch plan of the vector kernel
// (bsm_plan_vector answers it as a dry run, which is how tests/test_launch_plan.py pins every heuristic on the CPU),
// and the nnz-balanced row partition of the multi-GPU path.
#include <algorithm>
#include <cmath>
#include <string>

#include "bsm_internal.h"

namespace bsm {

static int pow2_ceil(int x)
{
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// lane shape for `n` columns starting at byte-aligned pointers
Shape pick_shape(uint32_t n, uint64_t ldb, uint64_t ldc, uint64_t col0, const void *b, const void *c, size_t s, bool prefer_wide,
                 uint64_t extra_ld)
{
    Shape sh;
    int V = (int)(16 / s);
    // lanes are 128-bit (2 x f64, 4 x f32) or one element wide; 2-element f32 lanes are not built (they only ever served column
    // counts that are even but not multiples of 4, at twice the kernel count)
    auto ok = [&](int v) {
        const uint64_t bytes = (uint64_t)v * s;
        if (s == 4 && v == 2) return false;
        return n % v == 0 && ldb % v == 0 && ldc % v == 0 && col0 % v == 0 && (extra_ld % v == 0) &&
               ((uintptr_t)b % bytes == 0) && ((uintptr_t)c % bytes == 0);
    };
    while (V > 1 && !ok(V)) V /= 2;
    int L = (int)(n / V);
    if (prefer_wide)
        while (L < 32 && V > 1) {
            V /= 2;
            if (s == 4 && V == 2) V = 1;
            L = (int)(n / V);
        }
    sh.V = V;
    if (L >= 32) {
        sh.G = 32;
        const int nt = (L + 31) / 32;
        sh.NT = nt <= 1 ? 1 : (nt <= 2 ? 2 : 4);
    } else {
        sh.G = pow2_ceil(L);
        sh.NT = 1;
    }
    return sh;
}

uint32_t column_tile(const bsm_tuning &tn, uint32_t n_total, int vmax, uint32_t tiles_max)
{
    uint32_t tile = tn.col_tile > 0 ? (uint32_t)tn.col_tile : n_total;
    tile = std::min<uint32_t>(tile, 32u * (uint32_t)vmax * tiles_max);   // widest shape one pass can hold in registers
    if (tile < n_total && tile % vmax) tile = std::max<uint32_t>(vmax, tile / vmax * vmax);
    return tile;
}

// ---- launch plan of the vector kernel for one column pass ------------------------------------------------
// Every default below comes from a same-box A/B sweep kept in profiles/ (tools/sweep.py); bsm_tuning overrides each.

// What the planner knows about the matrix and the device — no pointers, no CUDA calls, so the same code plans a
// launch for bsm_spmm and answers bsm_plan_vector (a dry run the CPU test-suite uses to pin the heuristics).

// Grouped lanes: fewer lanes per row than the 128-bit loads need -> 2 or 4 register tiles per lane and 32/G rows side
// by side, every lane group walking its own flat entry stream over a run of consecutive rows (one LDS of the staged A
// stream then feeds 32/G rows). Exists for full-width 128-bit shapes only. Returns true when `sh` was regrouped;
// `by_default` says the choice was the heuristic's, not the caller's.
//   defaults on regular rows (uneven rows stay on the warp-per-row stream, which balances them inside the warp;
//   3-D Laplacian, profiles/r1_sweep{u,v,w,y,z,aa}_l3d_*.jsonl):
//     one 128-bit tile per lane, 512-byte rows (x64 f64): 8 lanes x 4 tiles            4.54 -> 3.92 ms
//     256-byte rows (x32 f64, x64 f32), short rows: 16 -> 8 lanes x 2 tiles            4.28 -> 2.22 ms
//     128-byte rows (x16 f64), short rows: 8 -> 4 lanes x 2 tiles                      2.70 -> 1.37 ms
//   (the row-by-row walk of narrow shapes drains its gather window at every row end; rows of ~65 entries, the band
//   matrix x32 f32, are still faster row by row: 0.45 vs 0.72 ms)
static bool regroup_lanes(const MatrixFacts &m, const bsm_tuning &tn, uint32_t n, size_t s, bool allowed, Shape &sh, bool &by_default)
{
    const double mean = m.mean();
    int want_g = tn.lanes_per_row;
    by_default = false;
    if (want_g == 0 && sh.NT == 1 && tn.reg_flavour <= 0 && tn.warps_per_cta <= 0 && tn.prefer_wide_rows == 0 &&
        (double)m.max_row_nnz <= 4.0 * mean + 8.0) {
        if (sh.G == 32) want_g = 8;
        else if ((sh.G == 16 || sh.G == 8) && mean <= 32.0) want_g = sh.G / 2;
        by_default = want_g > 0;
    }
    if (!allowed || want_g <= 0 || want_g >= sh.G || sh.V * (int)s != 16 || n != (uint32_t)(sh.V * sh.G * sh.NT)) return false;
    const int nt = (int)(n / (uint32_t)(sh.V * want_g));
    if (!(want_g == 16 || want_g == 8 || want_g == 4) || !(nt == 2 || nt == 4) || n != (uint32_t)(sh.V * want_g * nt)) return false;
    sh.G = want_g;
    sh.NT = nt;
    return true;
}

// Register-budget flavour (index into the table of spmm_rows_inst.cuh; bsm_tuning.reg_flavour is this + 1).
//   full-width G == 32: several tiles per lane -> 3 CTAs x 8 warps per SM (4); one tile per lane -> one CTA of 24 warps,
//   window of 10 gathers (7); scalar A-stream reads in both (LDS.128 reads measured 5-8 % slower);
//   grouped lanes: one CTA of 24 warps (6) for >= 8 lanes by default, else 3 x 8 warps (4);
//   narrow one-tile shapes: 0, or 4 = scalar A-stream reads (the default for a row per lane);
//   the scatter variant exists for the default flavours only.
static int pick_row_flavour(const bsm_tuning &tn, const Shape &sh, bool wide_full, bool grouped, bool grouped_by_default, bool multi)
{
    const bool user_nw = tn.warps_per_cta > 0;
    int flavour = tn.reg_flavour > 0 ? std::min(tn.reg_flavour, 8) - 1 : (wide_full && !user_nw ? (sh.NT >= 2 ? 4 : 7) : (sh.G == 1 ? 4 : 0));
    if (!wide_full && !(sh.G < 32 && flavour == 4)) flavour = 0;
    if (grouped) flavour = (tn.reg_flavour == 7 || tn.reg_flavour == 8 || (grouped_by_default && tn.reg_flavour <= 0 && sh.G >= 8)) ? 6 : 4;
    // flavours the round-1 sweeps rejected are no longer built (spmm_rows_inst.cuh): deeper window -> 0, LDS.128 reads -> scalar
    if (flavour == 1) flavour = 0;
    if (flavour == 2 || flavour == 3) flavour = 4;
    if (flavour == 5) flavour = 6;
    if (flavour == 7 && sh.NT >= 2) flavour = 4;         // the deep window exists for one tile per lane only
    if (multi) flavour = 0;   // one scatter kernel per shape (512 threads, 1 CTA per SM)
    return flavour;
}

// Rows per warp inside a super-batch: the dominant row stride of a stencil-like matrix (so the warps of a CTA sweep
// adjacent grid lines), else one slice (a few slices for narrow shapes on large matrices: measured on band x 32).
static uint32_t pick_rows_per_warp(const MatrixFacts &m, const DeviceFacts &dev, const bsm_tuning &tn, const Shape &sh, uint32_t R, int nw, int resident)
{
    uint32_t P = R;
    if (tn.rows_per_warp <= 0 && sh.G < 32 && m.rows / ((uint64_t)nw * 4 * R) >= 4ull * dev.sm_count) P = 4 * R;
    if (tn.rows_per_warp > 0) {
        P = (uint32_t)tn.rows_per_warp;
    } else if (m.row_stride >= 2 * R) {
        // P = stride / m keeps warps w and w+m on adjacent lines. Among stride, stride/2, stride/4, ... pick the one that
        // wastes least to wave quantisation (rounds x rows per warp per round); a larger P wins unless a smaller one
        // saves more than 10 % (measured on 1/8 and 1/4 row blocks: profiles/r1_sweepk_l3d_n128_s8.jsonl — locality
        // beats balance). Too few rows for even one round per SM: plain slices.
        const uint64_t grid_est = (uint64_t)dev.sm_count * resident;
        double best_cost = 0.0;
        uint32_t best_p = 0;
        for (uint32_t cand = m.row_stride; cand >= 2 * R; cand /= 2) {
            const uint64_t supers = (m.rows + (uint64_t)nw * cand - 1) / ((uint64_t)nw * cand);
            const double cost = (double)((supers + grid_est - 1) / grid_est) * cand;
            if (best_p == 0 || cost < 0.90 * best_cost) {
                best_cost = cost;
                best_p = cand;
            }
            if (cand % 2) break;
        }
        P = best_p ? best_p : R;
        if (m.rows / ((uint64_t)nw * P) < (uint64_t)dev.sm_count) P = R;
    }
    // 32- and 64-byte output rows on a stencil matrix: half / a quarter of a line per warp (x8 f64 0.951 -> 0.927 ms, x4 f64
    // 0.660 -> 0.640 ms, profiles/r2_sweep_narrow_final_*.jsonl: these shapes are bound by L1 wavefronts, not by the sharing of
    // neighbouring lines between the warps of a CTA, and shorter runs balance better)
    if (tn.rows_per_warp <= 0 && sh.NT == 1 && (sh.G == 2 || sh.G == 4) && m.row_stride >= 2 * R && P == m.row_stride) P = std::max(R, P / (sh.G == 2 ? 4u : 2u));
    // flat-stream shapes take any P >= R (the last slice of a line may be short; the row_ptr windows are realigned in
    // the kernel); the row-by-row narrow shapes keep whole slices
    if (sh.G == 32 || sh.NT > 1) return std::max(R, P);
    return std::max(R, (P + R - 1) / R * R);
}


// Launch plan of the vector kernel for one pass of n columns. Pure host arithmetic.
int plan_vector_pass(const MatrixFacts &m, const DeviceFacts &dev, const bsm_tuning &tn, uint32_t n, const PassAlign &al, bool scatter,
                            bool multi, VectorPlan *out)
{
    const size_t s = dtype_size(m.dtype);
    const double mean = m.mean();
    const bool user_R = tn.rows_per_slice > 0, user_nw = tn.warps_per_cta > 0, user_stages = tn.stages > 0;
    const size_t smem_max = dev.smem_max;
    RowParams p{};   // geometry fields only (row_kernel_smem_bytes reads cap, R, stages)
    // second attempt = without grouped lanes, when their (always staged) slices do not fit shared memory
    for (bool allow_grouped = true;; allow_grouped = false) {
        Shape sh = pick_shape(n, al.ldb, al.ldc, al.col0, al.b, al.c, s, tn.prefer_wide_rows != 0);
        bool grouped_by_default = false;
        const bool grouped = regroup_lanes(m, tn, n, s, allow_grouped && !scatter, sh, grouped_by_default);
        const bool wide_full = sh.G == 32 && n == (uint32_t)(sh.V * sh.G * sh.NT);
        const uint32_t rpp = 32u / (uint32_t)sh.G;   // rows side by side in one warp
        const uint32_t rq = std::max(4u, rpp);       // slice granularity (rpp is a power of two)
        // 64-byte output rows (four 128-bit lanes, one tile) on short regular rows: every lane group walks ONE flat entry stream over
        // its run of rows instead of row by row, so the eight gather windows of a warp never drain at a row end (flavour 8;
        // 3-D Laplacian x8 f64 1.26 -> 0.96 ms, profiles/r2_sweep_flat_narrow_*.jsonl; two-lane shapes and long rows do not gain)
        const bool flat_narrow = !multi && !scatter && !grouped && sh.G == 4 && sh.NT == 1 && (size_t)sh.V * s == 16 &&
                                 (tn.reg_flavour == 9 || (tn.reg_flavour <= 0 && !user_nw && mean <= 16.0 && (double)m.max_row_nnz <= 4.0 * mean + 8.0));

        // rows per TMA slice: ~128 entries per bulk copy (~224 with one register tile per lane, r1_sweepi_*, and for
        // 8 lanes x 2 tiles); narrow shapes want several row passes per slice to amortise the slice bookkeeping
        uint32_t R;
        if (user_R) {
            R = (uint32_t)tn.rows_per_slice;
        } else {
            const double target = ((sh.G == 32 && sh.NT == 1) || (grouped && grouped_by_default && sh.NT == 2)) ? 224.0 : 128.0;
            R = (uint32_t)std::min<double>(256.0, std::max(1.0, target / std::max(1.0, mean)));
            if (sh.G < 32) R = std::max(R, 4u * rpp);
            // a flat stream wants ~56 entries per lane group and slice (seven window turns after one turn of prologue)
            if (flat_narrow) R = rpp * (uint32_t)std::min(16.0, std::max(4.0, std::ceil(56.0 / std::max(1.0, mean))));
        }
        R = std::max(rq, R / rq * rq);
        // a stencil-like matrix: the slice must divide the line length, or the rows per warp (a multiple of the
        // slice) stop matching the lines and the L1 sharing between the warps of a CTA is lost (measured on a
        // 5-entry-per-row stencil, line 256: 24-row slices -> P = 264: 12.9 ms; 16-row slices -> P = 256: see
        // profiles/r1_probe_near_diag.jsonl)
        // (a divisor down to half the target; failing that the last slice of every line is short — a line of 100
        // keeps 16-row slices, 6 x 16 + 4, rather than dropping to 4-row slices)
        if (!user_R && m.row_stride >= 2 * rq && m.row_stride % R) {
            uint32_t r2 = R;
            while (r2 > rq && 2 * r2 > R && m.row_stride % r2) r2 -= rq;
            if (m.row_stride % r2 == 0) R = r2;
        }

        int flavour = pick_row_flavour(tn, sh, wide_full, grouped, grouped_by_default, multi);
        if (flat_narrow) flavour = 8;
        // warps per CTA: what the flavour was compiled for; fewer on small matrices, so that no SM idles behind a
        // handful of fat super-batches
        const bool big_cta = (wide_full && (flavour == 5 || flavour == 6 || flavour == 7)) || (grouped && flavour == 6);
        const int max_warps = big_cta ? 24 : ((flavour >= 2 && (wide_full || grouped)) ? 8 : 16);
        int nw = user_nw ? std::min(tn.warps_per_cta, 24) : (big_cta ? 24 : 16);
        if (!user_nw) {
            const uint64_t rows_per_warp_min = big_cta ? R : rq;
            const int nw_floor = big_cta ? 3 : 2;
            while (nw > nw_floor && (uint64_t)nw * rows_per_warp_min * (uint64_t)dev.sm_count > m.rows) nw /= 2;
        }
        nw = std::min(nw, max_warps);

        // The stage must hold the entries of ANY R consecutive rows (+3 for the 16-byte aligned start, + slack: the
        // vectorised A-stream reads run up to two gather windows past the slice). Shrink, in this order, the ring
        // depth, the slice and the CTA until the rings fit: first under a soft limit that leaves most of the 228 KB
        // to L1 (where wide B rows live), then under the hardware limit. If even the smallest slice cannot be
        // staged, col_idx / values are read from global memory instead (unstaged variant).
        const int resident = ((wide_full || grouped) && (flavour == 2 || flavour == 4)) ? 3 : 1;   // CTAs per SM the flavour targets
        const size_t smem_soft = (size_t)n * s <= 64 ? smem_max : std::min<size_t>(smem_max, (160 * 1024) / resident);
        const uint32_t window = flavour == 7 ? 10u : (sh.NT >= 4 ? 2u : (sh.NT == 2 ? 4u : 8u)) * (flavour == 1 ? 2u : 1u);   // gathers in flight
        const uint32_t r_floor = sh.G < 32 ? std::max(rq, 2u * rpp) : rq;
        p.R = R;
        p.stages = user_stages ? (uint32_t)std::min(tn.stages, 8) : (grouped && grouped_by_default && sh.G >= 8 ? 2u : 3u);
        auto smem_now = [&]() {
            p.cap = (uint32_t)pad4((uint64_t)p.R * m.max_row_nnz + 3) + (uint32_t)pad4(2 * window) + 4;   // a multiple of 4: the stage arrays stay 16-byte aligned
            return row_kernel_smem_bytes(m.dtype, p, nw);
        };
        size_t smem = smem_now();
        while (smem > smem_soft && p.stages > 2 && !user_stages) { --p.stages; smem = smem_now(); }
        while (smem > smem_soft && p.R > r_floor && !user_R) { p.R = std::max(r_floor, p.R / 2 / rq * rq); smem = smem_now(); }
        while (smem > smem_soft && nw > 8 && !user_nw) { nw /= 2; smem = smem_now(); }
        while (smem > smem_max && p.stages > 1) { --p.stages; smem = smem_now(); }
        while (smem > smem_max && p.R > rq && !user_R) { p.R = std::max(rq, p.R / 2 / rq * rq); smem = smem_now(); }
        while (smem > smem_max && nw > 2 && !user_nw) { nw /= 2; smem = smem_now(); }
        if (smem > smem_max && grouped) continue;   // the grouped shapes have no unstaged variant: a warp per row instead
        if (smem > smem_max) {                      // rows too long to stage: unstaged variant (row_ptr windows only)
            flavour = -1;
            p.cap = 0;
            p.stages = user_stages ? (uint32_t)std::min(tn.stages, 8) : 3u;
            smem = row_kernel_smem_bytes(m.dtype, p, nw);
        }
        if (smem > smem_max) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_vector: slice ring does not fit shared memory");

        p.P = pick_rows_per_warp(m, dev, tn, sh, p.R, nw, resident);
        const uint64_t S = (uint64_t)nw * p.P;
        out->sh = sh;
        out->grouped = grouped;
        out->flavour = flavour;
        out->nw = nw;
        out->R = p.R;
        out->P = p.P;
        out->stages = p.stages;
        out->cap = p.cap;
        out->num_super = (uint32_t)((m.rows + S - 1) / S);
        out->smem = smem;
        out->resident = resident;
        return BSM_OK;
    }
}

}  // namespace bsm

using namespace bsm;

extern "C" {

// Dry run of the vector kernel's launch heuristics — pure host arithmetic, no device needed (the CPU test-suite pins
// the heuristics with it). Operands are assumed 16-byte aligned with ld = n rounded up to 16 bytes; `grid` assumes the
// occupancy the chosen flavour targets.
int bsm_plan_vector(int dtype, uint64_t rows, uint64_t nnz, uint64_t max_row_nnz, uint32_t row_stride, uint64_t n_cols,
                    const bsm_tuning *tuning, int sm_count, uint64_t smem_optin_bytes, bsm_launch_info *out)
{
    if (!out) return fail(BSM_ERR_INVALID_ARGUMENT, "plan_vector: null out");
    if (dtype != BSM_F32 && dtype != BSM_F64) return fail(BSM_ERR_DTYPE_MISMATCH, "plan_vector: dtype must be f32 or f64");
    if (sm_count <= 0 || smem_optin_bytes < 2048 || n_cols == 0 || n_cols >= 0xFFFFFFF0ull)
        return fail(BSM_ERR_INVALID_ARGUMENT, "plan_vector: bad device facts or column count");
    bsm_tuning tn{};
    if (tuning) tn = *tuning;
    const size_t s = dtype_size(dtype);
    const int vmax = (int)(16 / s);
    const uint32_t n_total = (uint32_t)n_cols;
    uint32_t tile = tn.col_tile > 0 ? (uint32_t)tn.col_tile : n_total;
    tile = std::min<uint32_t>(tile, 32u * vmax * 4u);
    if (tile < n_total && tile % vmax) tile = std::max<uint32_t>(vmax, tile / vmax * vmax);
    const MatrixFacts m{dtype, rows, nnz, max_row_nnz, row_stride};
    const DeviceFacts dev{sm_count, (size_t)smem_optin_bytes - 1024};
    const uint64_t ld = round_up(n_total, vmax);
    // the plan reported is that of the FIRST pass; `passes` counts them all, `capacity`... see bsm_launch_info
    uint32_t first_n = 0;
    int passes = 0;
    for (uint32_t col0 = 0, n = 0; col0 < n_total; col0 += n, ++passes) {
        n = fit_pass_width(std::min(tile, n_total - col0), vmax, [&](uint32_t w) { return pick_shape(w, ld, ld, col0, nullptr, nullptr, s, tn.prefer_wide_rows != 0); });
        if (n == 0) return fail(BSM_ERR_INVALID_ARGUMENT, "plan_vector: empty pass");
        if (!first_n) first_n = n;
    }
    VectorPlan plan;
    BSM_TRY(plan_vector_pass(m, dev, tn, first_n, PassAlign{ld, ld, 0, nullptr, nullptr}, false, false, &plan));
    *out = bsm_launch_info();
    out->algo = BSM_ALGO_VECTOR;
    out->vec_elems = plan.sh.V;
    out->lanes_per_row = plan.sh.G;
    out->reg_tiles = plan.sh.NT;
    out->block = plan.nw * 32;
    out->grid = (int)std::min<uint64_t>(plan.num_super, (uint64_t)sm_count * plan.resident);
    out->smem_bytes = (int)plan.smem;
    out->rows_per_slice = (int)plan.R;
    out->rows_per_warp = (int)plan.P;
    out->stages = (int)plan.stages;
    out->capacity = (int)plan.cap;
    out->reg_flavour = plan.flavour + 1;
    out->col_tile = (int)first_n;   // width of the first pass
    out->passes = passes;
    return BSM_OK;
}

// ---- sharding ------------------------------------------------------------------------------------
int bsm_partition_rows(const uint64_t *row_index, uint64_t rows, int parts, uint64_t *bounds)
{
    if (!row_index || !bounds || parts < 1) return fail(BSM_ERR_INVALID_ARGUMENT, "partition_rows: bad arguments");
    const uint64_t nnz = row_index[rows] - row_index[0];
    bounds[0] = 0;
    for (int p = 1; p < parts; ++p) {
        // first row whose start offset reaches p/parts of the entries (ties -> equal row counts)
        const uint64_t target = row_index[0] + (uint64_t)((__uint128_t)nnz * (unsigned)p / (unsigned)parts);
        uint64_t r;
        if (nnz == 0) {
            r = rows * (uint64_t)p / (uint64_t)parts;
        } else {
            r = (uint64_t)(std::lower_bound(row_index, row_index + rows + 1, target) - row_index);
            if (r > rows) r = rows;
        }
        bounds[p] = std::max(r, bounds[p - 1]);
    }
    bounds[parts] = rows;
    return BSM_OK;
}

}  // extern "C"
