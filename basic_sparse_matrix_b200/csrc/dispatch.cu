// dispatch.cu — THE HOT PATH's host side: which kernel family runs a product (vector CSR / merge-path / row-block)
// and with what geometry. Replaces the call Csr::mul_dense, /root/reference/src/sparse.rs:426-446.
#include <algorithm>
#include <cmath>
#include <string>

#include "bsm_internal.h"

namespace bsm {

static MatrixFacts facts_of(const bsm_csr *a) { return MatrixFacts{a->dtype, a->rows, a->nnz, a->max_row_nnz, a->row_stride}; }
static double mean_row_nnz(const bsm_csr *a) { return facts_of(a).mean(); }

// Scatter variant of the vector kernel: C is this rank's FULL result buffer, the rank's rows start at
// row_offset, and every row is also stored to `n_peers` further full buffers (peer GPUs over NVLink).
struct ScatterTargets {
    int n_peers = 0;
    void *peer_data[7] = {};
    uint64_t row_offset = 0;
};

static int spmm_vector(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning &tn, uint32_t flags, cudaStream_t stream,
                       const ScatterTargets *scatter = nullptr)
{
    const size_t s = dtype_size(a->dtype);
    const uint32_t n_total = (uint32_t)b->cols;
    const int vmax = (int)(16 / s);
    const uint32_t tile = column_tile(tn, n_total, vmax, 4);
    const bool multi = scatter && scatter->n_peers > 0;
    const MatrixFacts m = facts_of(a);
    const DeviceFacts dev{rt().sm_count, (size_t)rt().max_smem_optin - 1024};
    int passes = 0;
    launch_info() = bsm_launch_info();
    launch_info().algo = BSM_ALGO_VECTOR;
    for (uint32_t col0 = 0, n = 0; col0 < n_total; col0 += n, ++passes) {
        n = fit_pass_width(std::min(tile, n_total - col0), vmax,
                           [&](uint32_t w) { return pick_shape(w, b->ld, c->ld, col0, b->data, c->data, s, tn.prefer_wide_rows != 0); });
        if (passes == 0) launch_info().col_tile = (int)n;   // width of the first pass (what bsm_plan_vector reports too)
        const size_t c_off = ((scatter ? (size_t)scatter->row_offset * c->ld : 0) + (size_t)col0) * s;
        RowParams p{};
        p.row_ptr = a->row_ptr;
        p.col_idx = a->col_idx;
        p.vals = a->vals;
        p.B = (const char *)b->data + (size_t)col0 * s;
        p.C = (char *)c->data + c_off;
        p.rows = (uint32_t)a->rows;
        p.n = n;
        p.ldb = (uint32_t)b->ld;
        p.ldc = (uint32_t)c->ld;
        p.flags = flags;
        if (multi) {
            p.n_peers = (uint32_t)scatter->n_peers;
            for (int d = 0; d < scatter->n_peers; ++d) p.peers[d] = (char *)scatter->peer_data[d] + c_off;
        }
        VectorPlan plan;
        BSM_TRY(plan_vector_pass(m, dev, tn, n, PassAlign{b->ld, c->ld, col0, b->data, c->data}, scatter != nullptr, multi, &plan));
        const Shape sh = plan.sh;
        const int flavour = plan.flavour, nw = plan.nw;
        const size_t smem = plan.smem;
        p.R = plan.R;
        p.P = plan.P;
        p.stages = plan.stages;
        p.cap = plan.cap;
        p.num_super = plan.num_super;
        {
            const int block = nw * 32;
            int occ = 0;
            BSM_TRY(row_kernel_occupancy(a->dtype, sh, n, flavour, multi, block, smem, &occ));
            if (occ < 1) return fail(BSM_ERR_CUDA, "spmm_vector: kernel does not fit on an SM");
            const int ctas = tn.ctas_per_sm > 0 ? std::min(tn.ctas_per_sm, occ) : std::min(occ, 4);
            const int grid = (int)std::min<uint64_t>(p.num_super, (uint64_t)rt().sm_count * ctas);
            if (grid > 0) BSM_TRY(launch_spmm_rows(a->dtype, sh, p, flavour, multi, grid, block, smem, ctas, stream));
            launch_info().kernels += grid > 0;
            launch_info().vec_elems = sh.V;
            launch_info().lanes_per_row = sh.G;
            launch_info().reg_tiles = sh.NT;
            launch_info().grid = grid;
            launch_info().block = block;
            launch_info().smem_bytes = (int)smem;
            launch_info().rows_per_slice = (int)p.R;
            launch_info().rows_per_warp = (int)p.P;
            launch_info().reg_flavour = flavour + 1;
            launch_info().stages = (int)p.stages;
            launch_info().capacity = (int)p.cap;
        }
    }
    launch_info().passes = passes;
    return BSM_OK;
}

// merge-path caches of a handle live where the handle's arrays live (pool or cudaMalloc); all of one kind
static int cache_alloc(bsm_csr *a, void **p, size_t bytes, cudaStream_t stream)
{
    if (!a->part_rows && !a->carry_vals && !a->long_rows) a->cache_pooled = a->pooled && a->owns;
    if (a->cache_pooled) {   // stream-ordered: on the stream the kernels that use the cache run on
        BSM_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, stream));
        return BSM_OK;
    }
    BSM_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    return BSM_OK;
}

static void cache_free(bsm_csr *a, void *p, cudaStream_t stream)
{
    if (!p) return;
    if (a->cache_pooled)
        cudaFreeAsync(p, stream);
    else
        cudaFree(p);
}

static int ensure_partition(bsm_csr *a, uint32_t items, uint32_t num_chunks, cudaStream_t stream)
{
    if (a->part_rows && a->part_items == (int)items && a->part_chunks == num_chunks) return BSM_OK;
    cache_free(a, a->part_rows, stream);
    a->part_rows = nullptr;
    BSM_TRY(cache_alloc(a, (void **)&a->part_rows, ((size_t)num_chunks + 1) * 4, stream));
    BSM_TRY(launch_merge_partition(a->row_ptr, (uint32_t)a->rows, (uint32_t)a->nnz, items, num_chunks, a->part_rows, stream));
    launch_info().kernels += 1;
    a->part_items = (int)items;
    a->part_chunks = num_chunks;
    return BSM_OK;
}

static int spmm_merge(const bsm_csr *a_const, const bsm_dense *b, bsm_dense *c, const bsm_tuning &tn, uint32_t flags, cudaStream_t stream)
{
    bsm_csr *a = const_cast<bsm_csr *>(a_const);   // partition / carry caches live in the handle
    const size_t s = dtype_size(a->dtype);
    const uint32_t n_total = (uint32_t)b->cols;
    const int vmax = (int)(16 / s);
    const uint32_t tile = column_tile(tn, n_total, vmax, 4);
    const uint64_t total = a->rows + a->nnz;
    launch_info() = bsm_launch_info();
    launch_info().algo = BSM_ALGO_MERGE;
    int passes = 0;
    for (uint32_t col0 = 0, n = 0; col0 < n_total; col0 += n, ++passes) {
        n = fit_pass_width(std::min(tile, n_total - col0), vmax, [&](uint32_t w) {
            return pick_shape(w, b->ld, c->ld, col0, b->data, c->data, s, tn.prefer_wide_rows >= 0, round_up(w, vmax));
        });
        if (passes == 0) launch_info().col_tile = (int)n;
        const uint64_t ldcar = round_up(n, vmax);
        Shape sh = pick_shape(n, b->ld, c->ld, col0, b->data, c->data, s, tn.prefer_wide_rows >= 0, ldcar);
        // fewer lanes per chunk, 2 or 4 register tiles per lane (128-bit lanes, full-width shapes): one LDS.128 of the
        // staged A stream then feeds 32/G chunks. Defaults from the same-box A/B on R-MAT (profiles/r1_sweepx_rmat_*):
        // 512-byte rows 16 lanes x 2 tiles, 192 items (3.08 -> 3.01 ms); 256-byte rows 8 lanes x 2 tiles, 160 items
        // (2.08 -> 1.84 ms, r1_sweepy_rmat_f32)
        int want_g = tn.lanes_per_row;
        uint32_t auto_items = 0;
        if (want_g == 0 && tn.prefer_wide_rows == 0 && tn.merge_items <= 0 && tn.warps_per_cta <= 0) {
            if ((size_t)n * s == 512) { want_g = 16; auto_items = 192; }
            if ((size_t)n * s == 256) { want_g = 8; auto_items = 160; }
        }
        bool grouped = false;
        if (want_g > 0) {
            const Shape sv = pick_shape(n, b->ld, c->ld, col0, b->data, c->data, s, false, ldcar);
            const int g = want_g, nt = sv.V * (int)s == 16 ? (int)(n / (uint32_t)(sv.V * g)) : 0;
            if ((g == 16 || g == 8) && nt == 2 && n == (uint32_t)(sv.V * g * nt)) {   // the grouped merge shapes that are built
                sh.V = sv.V;
                sh.G = g;
                sh.NT = nt;
                grouped = true;
            }
        }
        const int nw = tn.warps_per_cta > 0 ? std::min(tn.warps_per_cta, 8) : 8;
        const int block = nw * 32;
        const uint32_t groups = (uint32_t)nw * (32u / sh.G);
        uint32_t items = tn.merge_items > 0 ? (uint32_t)tn.merge_items : (grouped && auto_items ? auto_items : (sh.G == 32 ? 384u : std::max(16u, 2048u / groups)));
        items = (uint32_t)round_up(items, 4);
        if (total + items >= 0xFFFFFFF0ull) return fail(BSM_ERR_INDEX_OVERFLOW, "spmm_merge: rows+nnz must fit u32");
        const uint32_t num_chunks = (uint32_t)((total + items - 1) / items);
        if (num_chunks == 0) continue;
        BSM_TRY(ensure_partition(a, items, num_chunks, stream));
        const size_t need_vals = (size_t)num_chunks * ldcar * s;
        if (a->carry_vals_bytes < need_vals) {
            cache_free(a, a->carry_vals, stream);
            a->carry_vals = nullptr;
            a->carry_vals_bytes = 0;
            BSM_TRY(cache_alloc(a, &a->carry_vals, need_vals, stream));
            a->carry_vals_bytes = need_vals;
        }
        // rows whose run of carries exceeds 64 chunks need more than 64*items entries: at most this many
        const uint64_t long_cap = a->nnz / (64ull * items) + 1;
        if (a->long_rows_cap < long_cap) {
            cache_free(a, a->long_rows, stream);
            a->long_rows = nullptr;
            a->long_rows_cap = 0;
            BSM_TRY(cache_alloc(a, (void **)&a->long_rows, (2 * long_cap + 1) * 4, stream));
            a->long_rows_cap = long_cap;
        }
        MergeParams p{};
        p.row_ptr = a->row_ptr;
        p.col_idx = a->col_idx;
        p.vals = a->vals;
        p.B = (const char *)b->data + (size_t)col0 * s;
        p.C = (char *)c->data + (size_t)col0 * s;
        p.part_rows = a->part_rows;
        p.carry_vals = a->carry_vals;
        p.long_rows = a->long_rows;
        p.long_count = a->long_rows + 2 * a->long_rows_cap;
        p.long_cap = (uint32_t)a->long_rows_cap;
        p.rows = (uint32_t)a->rows;
        p.nnz = (uint32_t)a->nnz;
        p.n = n;
        p.ldb = (uint32_t)b->ld;
        p.ldc = (uint32_t)c->ld;
        p.ldcar = (uint32_t)ldcar;
        p.items = items;
        p.num_chunks = num_chunks;
        p.flags = flags;
        const size_t smem = merge_kernel_smem_bytes(a->dtype, sh, block, items);
        if (smem > (size_t)rt().max_smem_optin - 1024) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_merge: items do not fit shared memory");
        int grid = 0;
        BSM_TRY(launch_spmm_merge(a->dtype, sh, p, block, smem, tn.ctas_per_sm, stream, &grid));
        int fix_launches = 0;
        BSM_TRY(launch_merge_fixup(a->dtype, p, stream, &fix_launches));
        launch_info().kernels += 1 + fix_launches;
        launch_info().vec_elems = sh.V;
        launch_info().lanes_per_row = sh.G;
        launch_info().reg_tiles = sh.NT;
        launch_info().grid = grid;
        launch_info().block = block;
        launch_info().smem_bytes = (int)smem;
        launch_info().merge_items = (int)items;
        launch_info().merge_chunks = (int)num_chunks;
    }
    launch_info().passes = passes;
    return BSM_OK;
}

// Row-block probe of a handle (cached): is every row a run of consecutive columns, and how many B rows would the
// row-block kernel load in all?
static int ensure_rowblock_probe(bsm_csr *a, cudaStream_t stream)
{
    if (a->rowblock_state) return BSM_OK;
    unsigned long long *d = nullptr, h[2] = {0, 0};
    BSM_CUDA(cudaMallocAsync(&d, 16, stream));
    BSM_CUDA(cudaMemsetAsync(d, 0, 16, stream));
    int st = launch_rowblock_probe(a->row_ptr, a->col_idx, a->rows, d, stream);
    if (st == BSM_OK && cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, stream) != cudaSuccess) st = fail(BSM_ERR_CUDA, "rowblock probe: copy failed");
    cudaFreeAsync(d, stream);
    BSM_TRY(st);
    BSM_CUDA(cudaStreamSynchronize(stream));
    a->rowblock_union = h[0];
    a->rowblock_state = h[1] ? 2 : 1;
    launch_info().kernels += 1;
    return BSM_OK;
}

// vector CSR for band-like matrices: blocks of kRowBlockRows consecutive rows share their B-row loads
static int spmm_rowblock(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning &tn, uint32_t flags, cudaStream_t stream)
{
    const size_t s = dtype_size(a->dtype);
    const uint32_t n_total = (uint32_t)b->cols;
    const int vmax = (int)(16 / s);
    const uint32_t tile = column_tile(tn, n_total, vmax, 1);   // one register tile per lane
    const int probe_kernels = launch_info().kernels;
    launch_info() = bsm_launch_info();
    launch_info().kernels = probe_kernels;
    launch_info().algo = BSM_ALGO_ROWBLOCK;
    int passes = 0;
    for (uint32_t col0 = 0, n = 0; col0 < n_total; col0 += n, ++passes) {
        n = std::min(tile, n_total - col0);
        Shape sh = pick_shape(n, b->ld, c->ld, col0, b->data, c->data, s, false);
        // the row-block kernel is built for 128-bit and one-element lanes, 4 to 32 lanes per row, one register tile per lane
        // (spmm_rowblock.cu): 2-element f32 lanes run as one-element lanes, narrower rows leave lanes idle, wider passes are cut
        if (s == 4 && sh.V == 2) sh.V = 1;
        uint32_t lanes = (n + (uint32_t)sh.V - 1) / (uint32_t)sh.V;
        if (lanes > 32) {
            lanes = 32;
            n = 32u * (uint32_t)sh.V;
        }
        sh.G = 4;
        while ((uint32_t)sh.G < lanes) sh.G *= 2;
        sh.NT = 1;
        // Two register tiles per lane on half as many lanes — the default for full-width 128-bit shapes of at least 8 lanes: one value
        // read (LDS) then feeds twice as many products, and the kernel is bound by the instructions it issues per product (band x32 f32
        // 0.283 -> 0.244 ms, x64 f64 1.15 -> 0.87 ms, x128 f32 0.96 -> 0.87 ms, bit-exact; profiles/r2_sweep_rowblock_tiles_*.jsonl).
        // bsm_tuning.lanes_per_row = the natural lane count keeps one tile per lane.
        if ((size_t)sh.V * s == 16 && n == (uint32_t)(sh.V * sh.G) && sh.G >= 8 && (tn.lanes_per_row <= 0 || tn.lanes_per_row * 2 == sh.G)) {
            sh.G /= 2;
            sh.NT = 2;
        }
        if (passes == 0) launch_info().col_tile = (int)n;
        RowBlockParams p{};
        p.row_ptr = a->row_ptr;
        p.col_idx = a->col_idx;
        p.vals = a->vals;
        p.B = (const char *)b->data + (size_t)col0 * s;
        p.C = (char *)c->data + (size_t)col0 * s;
        p.rows = (uint32_t)a->rows;
        p.n = n;
        p.ldb = (uint32_t)b->ld;
        p.ldc = (uint32_t)c->ld;
        p.flags = flags;
        int grid = 0, block = 0, smem = 0;
        // rows one lane group accumulates side by side: 8 when a warp holds one block (G = 32), 4 when it holds several
        // (more warps fit; band x32 f32 0.328 -> 0.296 ms, x128 f32 1.26 -> 1.15 ms the other way: r1_sweepag_band_*)
        int rb = tn.rows_per_slice == 4 || tn.rows_per_slice == 8 ? tn.rows_per_slice : (sh.G == 32 ? 8 : 4);
        BSM_TRY(launch_spmm_rowblock(a->dtype, sh, p, rb, a->max_row_nnz, rt().sm_count, (size_t)rt().max_smem_optin - 1024, stream, &grid, &block, &smem, &rb));
        launch_info().smem_bytes = smem;
        launch_info().kernels += grid > 0;
        launch_info().vec_elems = sh.V;
        launch_info().lanes_per_row = sh.G;
        launch_info().reg_tiles = sh.NT;
        launch_info().grid = grid;
        launch_info().block = block;
        launch_info().rows_per_slice = rb;
    }
    launch_info().passes = passes;
    return BSM_OK;
}

// Which kernel family runs a product (`requested` = bsm_algo of the caller, AUTO = the heuristics below).
static int choose_algo(const bsm_csr *a_const, uint64_t n_cols, int requested, int *algo, cudaStream_t stream)
{
    bsm_csr *a = const_cast<bsm_csr *>(a_const);   // the probe result is cached in the handle
    if (requested == BSM_ALGO_ROWBLOCK) {
        BSM_TRY(ensure_rowblock_probe(a, stream));
        if (a->rowblock_state != 1)
            return fail(BSM_ERR_NOT_SUPPORTED, "BSM_ALGO_ROWBLOCK: the matrix has a row whose stored columns are not a run of consecutive indices");
        *algo = BSM_ALGO_ROWBLOCK;
        return BSM_OK;
    }
    *algo = requested;
    if (requested == BSM_ALGO_VECTOR || requested == BSM_ALGO_MERGE) return BSM_OK;
    // csr_row_stats heuristic: the vector kernel serialises a row on one lane group, so one row far
    // above the mean (power-law hubs, the bench-as-written matrix) needs the nnz-balanced kernel
    const double mean = mean_row_nnz(a);
    *algo = BSM_ALGO_MERGE;
    if ((double)a->max_row_nnz > 64.0 + 8.0 * mean) return BSM_OK;
    // few, long rows (down to one giant row: a checksum vector, the bench-as-written matrix): fewer rows
    // than the vector kernel has warps, so only an entry-balanced split fills the machine
    if (a->rows < (uint64_t)rt().sm_count * 96 && a->max_row_nnz > 1024) return BSM_OK;
    *algo = BSM_ALGO_VECTOR;
    // band-like: long regular rows that are runs of consecutive columns, neighbouring rows sharing most of them
    // (at least 2x fewer B-row loads than entries) -> the row-block variant of the vector kernel
    // (output rows of at least 64 bytes: with fewer lanes per row the blocks' value reads are too scattered — SpMV on the
    // band measured 0.50 vs 0.11 ms)
    if (mean >= 16.0 && (double)a->max_row_nnz <= 4.0 * mean + 8.0 && a->rows >= (uint64_t)rt().sm_count * 64 &&
        n_cols * dtype_size(a->dtype) >= 64) {
        BSM_TRY(ensure_rowblock_probe(a, stream));
        if (a->rowblock_state == 1 && a->rowblock_union * 2 <= a->nnz) *algo = BSM_ALGO_ROWBLOCK;
    }
    return BSM_OK;
}

int csr_rows_are_runs(const bsm_csr *a_const, cudaStream_t stream, bool *runs)
{
    bsm_csr *a = const_cast<bsm_csr *>(a_const);   // the probe result is cached in the handle
    BSM_TRY(ensure_rowblock_probe(a, stream));
    *runs = a->rowblock_state == 1;
    return BSM_OK;
}

int resolve_algo(const bsm_csr *a, uint64_t n_cols, int requested, int *algo, cudaStream_t stream)
{
    return choose_algo(a, n_cols, requested, algo, stream);
}

int spmm_dispatch(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning *tuning, cudaStream_t stream)
{
    BSM_TRY(ensure_init());
    if (!a || !b || !c) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm: null handle");
    // src/sparse.rs:427-429
    if (a->cols != b->rows) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "spmm: A.cols != B.rows (MatErr::IncorrectDimensions)");
    if (c->rows != a->rows || c->cols != b->cols)
        return fail(BSM_ERR_INCORRECT_DIMENSIONS, "spmm: C must be A.rows x B.cols");
    if (a->dtype != b->dtype || a->dtype != c->dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "spmm: dtype mismatch");
    if (c->data == b->data && c->rows && c->cols) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm: C must not alias B");
    bsm_tuning tn{};
    if (tuning) tn = *tuning;
    uint32_t flags = tn.flags ? (tn.flags & 0x7FFFFFFFu) : BSM_TUNE_DEFAULT_FLAGS;
    launch_info() = bsm_launch_info();
    if (a->rows == 0 || b->cols == 0) return BSM_OK;
    int algo = BSM_ALGO_VECTOR;
    BSM_TRY(choose_algo(a, b->cols, tn.algo, &algo, stream));
    if (algo == BSM_ALGO_MERGE) return spmm_merge(a, b, c, tn, flags, stream);
    if (algo == BSM_ALGO_ROWBLOCK) {
        const int st = spmm_rowblock(a, b, c, tn, flags, stream);
        // chosen by the heuristic but a warp block's values do not fit shared memory: the plain vector kernel
        if (st != BSM_ERR_NOT_SUPPORTED || tn.algo == BSM_ALGO_ROWBLOCK) return st;
    }
    return spmm_vector(a, b, c, tn, flags, stream);
}

}  // namespace bsm

using namespace bsm;

extern "C" {

// ---- hot path ------------------------------------------------------------------------------------
int bsm_spmm(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, int algo)
{
    bsm_tuning tn{};
    tn.algo = algo;
    return spmm_dispatch(a, b, c, &tn, rt().stream);
}
int bsm_spmm_tuned(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning *tuning)
{
    return spmm_dispatch(a, b, c, tuning, rt().stream);
}
int bsm_spmm_scatter(const bsm_csr *a, const bsm_dense *b, bsm_dense *const *c_full, int ndest, uint64_t row_offset, int algo)
{
    BSM_TRY(ensure_init());
    if (!a || !b || !c_full || ndest < 1 || ndest > 8) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_scatter: bad arguments (1..8 destinations)");
    if (a->cols != b->rows) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "spmm_scatter: A.cols != B.rows (MatErr::IncorrectDimensions)");
    ScatterTargets st;
    st.row_offset = row_offset;
    for (int d = 0; d < ndest; ++d) {
        const bsm_dense *c = c_full[d];
        if (!c) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_scatter: null destination");
        if (c->dtype != a->dtype || b->dtype != a->dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "spmm_scatter: dtype mismatch");
        if (c->cols != b->cols || c->rows < row_offset + a->rows || c->ld != c_full[0]->ld)
            return fail(BSM_ERR_INCORRECT_DIMENSIONS, "spmm_scatter: every destination must be (>= row_offset + A.rows) x B.cols with one leading dimension");
        if (c->data == b->data) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_scatter: a destination aliases B");
        // the lane shape (up to 128-bit stores) is chosen from destination 0; every other destination is written with the same
        // vectors at the same offsets, so it needs the same alignment
        if (d > 0 && (((uintptr_t)c->data ^ (uintptr_t)c_full[0]->data) & 15)) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_scatter: every destination must be aligned like the first (mod 16 bytes)");
        if (d > 0) st.peer_data[st.n_peers++] = c->data;
    }
    // a row may be summed by one lane group only (its row is written, never read back): vector kernel
    int chosen = algo;
    if (algo == BSM_ALGO_AUTO) BSM_TRY(choose_algo(a, b->cols, algo, &chosen, rt().stream));
    if (chosen == BSM_ALGO_MERGE)
        return fail(BSM_ERR_NOT_SUPPORTED, "spmm_scatter: the merge-path kernel revisits C rows (fix-up) and cannot scatter; "
                                           "use bsm_spmm + bsm_allgather_rows for power-law matrices");
    launch_info() = bsm_launch_info();
    if (a->rows == 0 || b->cols == 0) return BSM_OK;
    bsm_tuning tn{};
    return spmm_vector(a, b, c_full[0], tn, BSM_TUNE_DEFAULT_FLAGS, rt().stream, &st);
}

int bsm_last_launch_info(bsm_launch_info *info)
{
    if (!info) return fail(BSM_ERR_INVALID_ARGUMENT, "last_launch_info: null");
    *info = launch_info();
    return BSM_OK;
}

}  // extern "C"
