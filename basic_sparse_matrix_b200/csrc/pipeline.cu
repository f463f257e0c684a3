// pipeline.cu — host-to-host entry points: the literal reference call Csr::mul_dense -> zero-dropped Csr
// (src/sparse.rs:426-446 incl. the result construction 442 -> 222-233 -> 206-219), its dense-result twin,
// Csr::mul_vector (468-482) and the residual norms of BASELINE config 5.
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "bsm_internal.h"

namespace bsm {

int dense_to_csr_impl(const bsm_dense *d, bsm_csr **out)
{
    BSM_TRY(ensure_init());
    if (!d || !out) return fail(BSM_ERR_INVALID_ARGUMENT, "dense_to_csr: null argument");
    cudaStream_t sm = rt().stream;
    uint32_t *counts = nullptr;
    unsigned long long *total = nullptr;
    bsm_csr *r = nullptr;
    int st = [&]() -> int {
        BSM_TRY(tmp_alloc((void **)&counts, (pad4(d->rows + 1) + 4) * 4));
        BSM_TRY(tmp_alloc((void **)&total, 8));
        BSM_CUDA(cudaMemsetAsync(total, 0, 8, sm));
        BSM_CUDA(cudaMemsetAsync(counts, 0, (pad4(d->rows + 1) + 4) * 4, sm));
        BSM_TRY(launch_count_nonzero(d->dtype, d->data, d->rows, d->cols, d->ld, counts, total, sm));
        BSM_TRY(exclusive_scan_u32(counts, counts, d->rows + 1, sm));   // row_index incl. the finalise() tail
        unsigned long long h = 0;
        BSM_CUDA(cudaMemcpyAsync(&h, total, 8, cudaMemcpyDeviceToHost, sm));
        BSM_CUDA(cudaStreamSynchronize(sm));
        BSM_TRY(alloc_csr(d->dtype, d->rows, d->cols, h, &r));
        BSM_CUDA(cudaMemcpyAsync(r->row_ptr, counts, (d->rows + 1) * 4, cudaMemcpyDeviceToDevice, sm));
        BSM_TRY(launch_scatter_nonzero(d->dtype, d->data, d->rows, d->cols, d->ld, r->row_ptr, r->vals, r->col_idx, sm));
        r->max_row_nnz = d->cols;
        BSM_CUDA(cudaStreamSynchronize(sm));
        return BSM_OK;
    }();
    tmp_free(counts);
    tmp_free(total);
    if (st != BSM_OK) {
        if (r) bsm_csr_free(r);
        return st;
    }
    *out = r;
    return BSM_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// The host-to-host product, pipelined: ONE routine, two result forms.
//
//   Csr result   = the literal reference call: Csr::mul_dense(&self, &Dense) -> Csr (sparse.rs:426-446), every output pushed
//                  through insert (zero-drop, 229) and finalise (206-219); arrays in the reference layout (usize indices);
//   Dense result = the same product as host Dense columns (no zero-drop), for callers that want the dense matrix.
//
// Stages, overlapped on three streams (PCIe is full duplex):
//   s_in   B travels host -> device in CHUNKS OF ROWS (n column pieces per chunk; the reference's Dense is column-major
//          Vec<Vec<T>>, dense.rs:5-9), only the window of rows A references;
//   s_mm   chunk transposes (column-major -> row-major), then per BLOCK OF OUTPUT ROWS: SpMM (as soon as the chunks its
//          columns reach have landed: a banded / stencil row block reads a window of B), and the result construction of
//          the block — count -> scan -> scatter (values + usize columns) + its row_index piece, or the transpose back to
//          column-major for the dense form;
//   s_out  the block's result travels device -> host while the next blocks are computed.
// The host blocks once per row block (it must know the block's entry count to size the copy) — one block behind the
// compute, so the GPU never waits for it.
// ------------------------------------------------------------------------------------------------------------------
namespace {

// device-side busy time of the pipeline's phases, measured with event pairs when the phase timers are on
struct DevTimeline {
    struct Span {
        int phase;
        cudaEvent_t a, b;
    };
    std::vector<Span> spans;
    bool on = phase_timers_on();
    void begin(int phase, cudaStream_t st)
    {
        if (!on) return;
        Span sp{phase, nullptr, nullptr};
        cudaEventCreate(&sp.a);
        cudaEventCreate(&sp.b);
        cudaEventRecord(sp.a, st);
        spans.push_back(sp);
    }
    void end(cudaStream_t st)
    {
        if (on && !spans.empty()) cudaEventRecord(spans.back().b, st);
    }
    void finish()   // after the streams have been synchronised
    {
        for (Span &sp : spans) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) phase_add(sp.phase, ms * 1e-3);
            cudaEventDestroy(sp.a);
            cudaEventDestroy(sp.b);
        }
        spans.clear();
    }
};

// The column indices of the result travel as ROW MASKS, not as usize values. The result of Csr x Dense is dense unless a sum
// happens to be exactly zero, so the returned Csr spends 8 bytes per entry on columns that say "0 .. n-1" in nearly every row —
// half of the device -> host traffic of the literal call, which is what bounds it. The device sends one keep-bit per output
// (count_nonzero_kernel writes them while it counts: n/8 bytes per row instead of 8 n), and a few host threads expand the masks of a
// row block into the caller's col_index while the values of the next blocks travel: a full row is one copy of the pattern
// 0 .. n-1, any other row is walked bit by bit. Same arrays, same contents as the reference's insert / finalise produce.
// BSM_PIPE_EXPAND_THREADS = number of threads (default 8, at most half of the hardware threads; 0 = the device writes usize
// columns and they are copied, as before). Measured on the headline product (34.5 GB of result): 717-772 ms -> 530-600 ms; the
// call is then bound by the host -> device copy of B, which the expansion's memory traffic slows from 350 to 500 ms.
class ColumnExpand {
    struct Job {
        cudaEvent_t landed;            // the block's masks and row_index piece are in host memory
        const uint64_t *masks;         // [rows][words]
        const uint64_t *row_index;     // absolute offsets of the rows into col_index
        uint64_t *col_index;
        uint64_t rows, n, words;
    };
    std::vector<std::thread> threads_;
    std::deque<Job> jobs_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool closing_ = false;
    int want_;
    int device_ = 0;
    std::vector<uint64_t> pattern_;   // 0 .. n-1
    bool nt_ = true;                  // non-temporal stores for full rows: no read-for-ownership of 17 GB nobody reads soon (BSM_PIPE_EXPAND_NT=0: plain)

    void fill_row(uint64_t *dst, uint64_t n) const
    {
#if defined(__x86_64__)
        if (nt_) {
            const uint64_t *pat = pattern_.data();
            uint64_t i = 0;
            if ((uintptr_t)dst & 15) {
                _mm_stream_si64(reinterpret_cast<long long *>(dst), 0);
                i = 1;
            }
            for (; i + 1 < n; i += 2) _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i *>(pat + i)));
            if (i < n) _mm_stream_si64(reinterpret_cast<long long *>(dst + i), (long long)i);
            return;
        }
#endif
        memcpy(dst, pattern_.data(), n * 8);
    }

    void work()
    {
        cudaSetDevice(device_);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return closing_ || !jobs_.empty(); });
                if (jobs_.empty()) return;
                j = jobs_.front();
                jobs_.pop_front();
            }
            cudaEventSynchronize(j.landed);
            const uint64_t full_tail = (j.n & 63) ? ((1ull << (j.n & 63)) - 1ull) : ~0ull;
            for (uint64_t r = 0; r < j.rows; ++r) {
                const uint64_t *m = j.masks + r * j.words;
                uint64_t *dst = j.col_index + j.row_index[r];
                bool full = m[j.words - 1] == full_tail;
                for (uint64_t w = 0; full && w + 1 < j.words; ++w) full = m[w] == ~0ull;
                if (full) {
                    fill_row(dst, j.n);
                    continue;
                }
                for (uint64_t w = 0; w < j.words; ++w)
                    for (uint64_t bits = m[w]; bits; bits &= bits - 1) *dst++ = w * 64 + (uint64_t)__builtin_ctzll(bits);
            }
        }
    }

public:
    ColumnExpand()
    {
        const char *e = getenv("BSM_PIPE_EXPAND_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        want_ = (e && *e) ? atoi(e) : std::min(8, std::max(1, hw / 2));
        if (want_ < 0) want_ = 0;
        const char *nt = getenv("BSM_PIPE_EXPAND_NT");   // 0 = plain stores (same-box A/B: 634-659 ms against 596-604 ms with non-temporal ones)
        nt_ = !(nt && *nt == '0');
        cudaGetDevice(&device_);
    }
    bool enabled() const { return want_ > 0; }
    // rows of one block; returns at once, the rows are split over the threads
    void expand(cudaEvent_t landed, const uint64_t *masks, const uint64_t *row_index, uint64_t *col_index, uint64_t rows, uint64_t n)
    {
        if (rows == 0 || n == 0) return;
        const uint64_t words = (n + 63) / 64;
        if (pattern_.size() != n) {
            pattern_.resize(n);
            for (uint64_t i = 0; i < n; ++i) pattern_[i] = i;
        }
        const uint64_t parts = std::min<uint64_t>((uint64_t)want_ * 2, std::max<uint64_t>(1, rows * n / (1u << 20)));
        const uint64_t step = (rows + parts - 1) / parts;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (uint64_t r = 0; r < rows; r += step)
                jobs_.push_back(Job{landed, masks + r * words, row_index + r, col_index, std::min(step, rows - r), n, words});
        }
        if (threads_.empty()) {
            try {
                for (int t = 0; t < want_; ++t) threads_.emplace_back([this] { work(); });
            } catch (...) {   // no thread could be started (resource limits): nothing may throw through the C ABI
            }
        }
        if (threads_.empty()) {   // expand on the calling thread: slower, same result
            {
                std::lock_guard<std::mutex> lk(mu_);
                closing_ = true;
            }
            work();
            closing_ = false;
            return;
        }
        cv_.notify_all();
    }
    // everything issued so far is done when this returns
    void finish()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            closing_ = true;
        }
        cv_.notify_all();
        for (std::thread &t : threads_) t.join();
        threads_.clear();
        closing_ = false;
    }
    ~ColumnExpand() { finish(); }
};

template <typename T> struct HostProduct {
    int dtype;
    uint64_t rows, cols, nnz;
    const T *v;
    const uint64_t *col_index, *row_index;
    uint64_t row_index_len, rhs_rows, rhs_cols;
    const T *const *rhs_col_ptrs;
    int algo;
    // dense result (column-major host Dense), or null for the Csr result
    T *const *out_col_ptrs;
    // Csr result: caller-provided arrays holding up to `capacity` entries (out_row_index: rows + 1)
    uint64_t capacity;
    T *out_v;
    uint64_t *out_col_index, *out_row_index, *out_nnz;
};

// Granularity. A block of output rows / a chunk of B rows moves over PCIe as n column pieces (the reference's Dense is one Vec
// per column), so the pieces must stay large: 1 MB pieces measured 30 GB/s host->device against 50 GB/s for contiguous
// 256 MB copies (profiles/r2_e2e_phases.md). Blocks and chunks are therefore at most kMaxBytes of row-major data and at least
// kMinPiece bytes per column piece, and in between an eighth of the operand so that a mid-size product still overlaps its copies.
constexpr uint64_t kMaxBytes = 1ull << 30;
constexpr uint64_t kMinPiece = 2ull << 20;
static uint64_t pipe_rows(uint64_t total_rows, uint64_t row_bytes, uint64_t elem_bytes, uint64_t override_bytes)
{
    if (override_bytes) return std::max<uint64_t>(1, override_bytes / row_bytes);
    const uint64_t hi = std::max<uint64_t>(1, kMaxBytes / row_bytes), lo = kMinPiece / elem_bytes;
    return std::min(hi, std::max(lo, total_rows / 8));
}
// BSM_PIPE_BLOCK_BYTES / BSM_PIPE_CHUNK_BYTES override the two (the tests shrink them to drive many blocks and chunks
// through small matrices)
static uint64_t env_bytes(const char *name)
{
    const char *e = getenv(name);
    return (e && *e) ? (uint64_t)strtoull(e, nullptr, 10) : 0;
}

template <typename T> int host_product(const HostProduct<T> &q)
{
    const bool want_csr = q.out_col_ptrs == nullptr;
    if (q.cols != q.rhs_rows) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "mul_dense: A.cols != rhs.rows (MatErr::IncorrectDimensions)");
    if (q.rhs_cols && !q.rhs_col_ptrs) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense: null column pointers");
    if (want_csr && (!q.out_nnz || !q.out_row_index || (q.capacity && (!q.out_v || !q.out_col_index))))
        return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense: null result arrays");
    BSM_TRY(ensure_init());
    PhaseScope total(PH_TOTAL);
    PoolScope pool;
    bsm_csr *a = nullptr;
    BSM_TRY(csr_upload<T>(q.dtype, q.rows, q.cols, q.nnz, q.v, q.col_index, q.row_index, q.row_index_len, &a));
    const uint64_t rows = q.rows, n = q.rhs_cols, s = sizeof(T);
    if (rows == 0 || n == 0) {
        bsm_csr_free(a);
        if (want_csr) {
            for (uint64_t r = 0; r <= rows; ++r) q.out_row_index[r] = 0;   // Csr::new + finalise: rows+1 zeros
            *q.out_nnz = 0;
        }
        return BSM_OK;
    }
    for (uint64_t c = 0; c < n; ++c)
        if (!q.rhs_col_ptrs[c] || (!want_csr && !q.out_col_ptrs[c])) {
            bsm_csr_free(a);
            return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense: null column");
        }

    PipelineStreams *ps = nullptr;
    cudaStream_t lib = rt().stream;
    DevTimeline tl;
    ColumnExpand expand;
    std::vector<cudaEvent_t> ev_landed;   // per row block: its masks and row_index piece have reached the host
    // geometry
    const uint64_t ld = default_ld(n, q.dtype);
    uint64_t rb = std::max<uint64_t>(4, pipe_rows(rows, ld * s, s, env_bytes("BSM_PIPE_BLOCK_BYTES")) / 4 * 4);       // rows per block (multiple of 4: the row_ptr
    if (a->row_stride && rb > a->row_stride) rb = rb / a->row_stride * a->row_stride;   // window of a view stays 16-byte
    if (rb % 4) rb = (rb + 3) / 4 * 4;                                           // aligned); whole stencil lines
    rb = std::min(rb, (rows + 3) / 4 * 4);
    if ((rows + rb - 1) / rb > 4096) rb = ((rows + 4095) / 4096 + 3) / 4 * 4;
    const uint32_t nblocks = (uint32_t)((rows + rb - 1) / rb);
    const uint64_t cb = std::min<uint64_t>(round_up(q.rhs_rows, 32), std::max<uint64_t>(32, pipe_rows(q.rhs_rows, n * s, s, env_bytes("BSM_PIPE_CHUNK_BYTES")) / 32 * 32));   // B rows per chunk
    // window of B rows A references (a rank's row block of a banded / stencil matrix reads a window of B, not all of it)
    const bool has_entries = a->nnz != 0;
    const uint64_t win_lo = has_entries ? (uint64_t)a->col_min / 32 * 32 : 0;
    const uint64_t win_hi = has_entries ? (uint64_t)a->col_max + 1 : 0;   // exclusive
    const uint64_t win_rows = win_hi - win_lo;

    T *bwin = nullptr, *cdev = nullptr, *stage_in[2] = {}, *stage_out[2] = {}, *vals_st[2] = {};
    uint64_t *cols_st[2] = {}, *rp64[2] = {}, *masks_st[2] = {}, *h_masks = nullptr;
    const uint64_t mwords = (n + 63) / 64;   // keep-bit words per result row
    // columns travel as row masks, expanded by host threads — for results large enough to pay for starting the threads (a small
    // product is launch-bound: the reference's own bench shape, 10^4 outputs, went from 1.9 to 3.5 ms with them)
    const uint64_t min_entries = getenv("BSM_PIPE_EXPAND_MIN_ENTRIES") ? env_bytes("BSM_PIPE_EXPAND_MIN_ENTRIES") : (4ull << 20);
    const bool by_masks = want_csr && expand.enabled() && rows * n >= min_entries;
    uint32_t *counts[2] = {}, *blk_range = nullptr;
    unsigned long long *tot = nullptr, *h_tot = nullptr;
    uint32_t *h_range = nullptr;
    bool streams_touched = false;

    int st = [&]() -> int {
        BSM_TRY(pipeline_streams(&ps));
        cudaStream_t s_in = ps->in, s_mm = ps->mm, s_out = ps->out, s_meta = ps->meta;
        cudaEvent_t ev_ready = ps->ev[0], *ev_in = &ps->ev[1], *ev_in_free = &ps->ev[3], *ev_c = &ps->ev[5], *ev_out_free = &ps->ev[7];
        // which kernel family: decided once, on the whole matrix
        int algo = BSM_ALGO_VECTOR;
        BSM_TRY(resolve_algo(a, n, q.algo, &algo, lib));
        const bool per_block = algo != BSM_ALGO_MERGE && nblocks > 1;   // the merge-path partition is per matrix: one SpMM over all rows
        // pinned scratch for the per-block entry counts and the per-block column ranges
        void *pin = nullptr;
        const size_t small = round_up(((size_t)nblocks + 1) * 8 + (size_t)nblocks * 8, 256);
        BSM_TRY(pinned_scratch(&pin, small + (by_masks ? (size_t)rows * mwords * 8 : 0)));
        h_tot = (unsigned long long *)pin;
        h_range = (uint32_t *)(h_tot + nblocks + 1);
        h_masks = (uint64_t *)((char *)pin + small);
        if (by_masks) {
            ev_landed.resize(nblocks, nullptr);
            for (uint32_t k = 0; k < nblocks; ++k) BSM_CUDA(cudaEventCreateWithFlags(&ev_landed[k], cudaEventDisableTiming));
        }
        // device buffers (stream-ordered pool on the library stream; the pipeline streams wait for ev_ready)
        BSM_TRY(tmp_alloc((void **)&bwin, std::max<uint64_t>(win_rows, 1) * ld * s + 16));
        BSM_TRY(tmp_alloc((void **)&cdev, rows * ld * s + 16));
        for (int i = 0; i < 2; ++i) {
            BSM_TRY(tmp_alloc((void **)&stage_in[i], cb * n * s));
            if (want_csr) {
                BSM_TRY(tmp_alloc((void **)&counts[i], (pad4(rb + 1) + 4) * 4));
                BSM_TRY(tmp_alloc((void **)&vals_st[i], rb * n * s));
                if (by_masks)
                    BSM_TRY(tmp_alloc((void **)&masks_st[i], rb * mwords * 8));
                else
                    BSM_TRY(tmp_alloc((void **)&cols_st[i], rb * n * 8));
                BSM_TRY(tmp_alloc((void **)&rp64[i], rb * 8));
            } else {
                BSM_TRY(tmp_alloc((void **)&stage_out[i], rb * n * s));
            }
        }
        if (ld != n) BSM_CUDA(cudaMemsetAsync(bwin, 0, std::max<uint64_t>(win_rows, 1) * ld * s, lib));   // defined padding columns
        if (want_csr) {
            BSM_TRY(tmp_alloc((void **)&tot, ((size_t)nblocks + 1) * 8));
            BSM_CUDA(cudaMemsetAsync(tot, 0, 8, lib));
            h_tot[0] = 0;
        }
        // per-block column ranges -> which chunk of B a row block has to wait for
        if (per_block && has_entries) {
            BSM_TRY(tmp_alloc((void **)&blk_range, (size_t)nblocks * 8));
            for (uint32_t k = 0; k < nblocks; ++k) {
                h_range[2 * k] = 0xFFFFFFFFu;
                h_range[2 * k + 1] = 0;
            }
            BSM_CUDA(cudaMemcpyAsync(blk_range, h_range, (size_t)nblocks * 8, cudaMemcpyHostToDevice, lib));
            BSM_TRY(launch_block_col_range(a->row_ptr, a->col_idx, rows, rb, nblocks, blk_range, lib));
            BSM_CUDA(cudaMemcpyAsync(h_range, blk_range, (size_t)nblocks * 8, cudaMemcpyDeviceToHost, lib));
            BSM_CUDA(cudaStreamSynchronize(lib));
        }
        BSM_CUDA(cudaEventRecord(ev_ready, lib));
        BSM_CUDA(cudaStreamWaitEvent(s_in, ev_ready, 0));
        BSM_CUDA(cudaStreamWaitEvent(s_mm, ev_ready, 0));
        BSM_CUDA(cudaStreamWaitEvent(s_out, ev_ready, 0));
        BSM_CUDA(cudaStreamWaitEvent(s_meta, ev_ready, 0));
        streams_touched = true;

        bsm_dense bd, cd;   // B as the kernels see it: row j at bwin + (j - win_lo) * ld
        bd.dtype = cd.dtype = q.dtype;
        bd.rows = q.rhs_rows;
        cd.rows = rows;
        bd.cols = cd.cols = n;
        bd.ld = cd.ld = ld;
        bd.data = (char *)bwin - win_lo * ld * s;
        cd.data = cdev;
        bd.owns = cd.owns = false;
        bsm_tuning tn{};
        tn.algo = algo;

        // ---- B chunks -------------------------------------------------------------------------------------------
        const uint64_t chunk_lo = win_lo / cb, chunk_end = has_entries ? (win_hi + cb - 1) / cb : chunk_lo;
        uint64_t next_chunk = chunk_lo, chunks_sent = 0;
        auto enqueue_chunk = [&](uint64_t j) -> int {
            const int i = (int)(chunks_sent & 1);
            const uint64_t p0 = std::max(j * cb, win_lo), p1 = std::min((j + 1) * cb, win_hi), pr = p1 - p0;
            if (chunks_sent >= 2) BSM_CUDA(cudaStreamWaitEvent(s_in, ev_in_free[i], 0));
            tl.begin(PH_B_H2D, s_in);
            for (uint64_t c = 0; c < n; ++c)
                BSM_CUDA(cudaMemcpyAsync(stage_in[i] + c * pr, q.rhs_col_ptrs[c] + p0, pr * s, cudaMemcpyHostToDevice, s_in));
            tl.end(s_in);
            BSM_CUDA(cudaEventRecord(ev_in[i], s_in));
            BSM_CUDA(cudaStreamWaitEvent(s_mm, ev_in[i], 0));
            tl.begin(PH_B_TRANSPOSE, s_mm);
            BSM_TRY(launch_transpose_cm2rm(q.dtype, stage_in[i], (T *)bwin + (p0 - win_lo) * ld, pr, n, ld, s_mm));
            tl.end(s_mm);
            BSM_CUDA(cudaEventRecord(ev_in_free[i], s_mm));
            ++chunks_sent;
            return BSM_OK;
        };

        // ---- result of row block k leaves the device ------------------------------------------------------------------
        auto drain = [&](uint32_t k) -> int {
            const int o = (int)(k & 1);
            const uint64_t r0 = (uint64_t)k * rb, rbk = std::min(rb, rows - r0);
            if (want_csr) {
                {
                    PhaseScope w(PH_WAIT);
                    BSM_CUDA(cudaEventSynchronize(ev_c[o]));   // h_tot[k+1] has landed
                }
                const uint64_t base = h_tot[k], nk = h_tot[k + 1] - base;
                if (base + nk > q.capacity)
                    return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense: the result arrays are too small (capacity " + std::to_string(q.capacity) +
                                                              " entries, need at least " + std::to_string(base + nk) + ")");
                tl.begin(PH_D2H, s_out);
                if (by_masks) {   // (masks and row_index piece are on their way already, see the block loop)
                    if (nk) {
                        BSM_CUDA(cudaMemcpyAsync(q.out_v + base, vals_st[o], nk * s, cudaMemcpyDeviceToHost, s_out));
                        expand.expand(ev_landed[k], h_masks + r0 * mwords, q.out_row_index + r0, q.out_col_index, rbk, n);
                    }
                } else {
                    if (nk) {
                        BSM_CUDA(cudaMemcpyAsync(q.out_v + base, vals_st[o], nk * s, cudaMemcpyDeviceToHost, s_out));
                        BSM_CUDA(cudaMemcpyAsync(q.out_col_index + base, cols_st[o], nk * 8, cudaMemcpyDeviceToHost, s_out));
                    }
                    BSM_CUDA(cudaMemcpyAsync(q.out_row_index + r0, rp64[o], rbk * 8, cudaMemcpyDeviceToHost, s_out));
                }
                tl.end(s_out);
            } else {
                BSM_CUDA(cudaStreamWaitEvent(s_out, ev_c[o], 0));
                tl.begin(PH_D2H, s_out);
                for (uint64_t c = 0; c < n; ++c)
                    BSM_CUDA(cudaMemcpyAsync(q.out_col_ptrs[c] + r0, stage_out[o] + c * rbk, rbk * s, cudaMemcpyDeviceToHost, s_out));
                tl.end(s_out);
            }
            BSM_CUDA(cudaEventRecord(ev_out_free[o], s_out));
            return BSM_OK;
        };

        // ---- row blocks --------------------------------------------------------------------------------------------
        uint64_t need_hi = 0;   // running maximum of the columns the blocks so far reach (exclusive)
        for (uint32_t k = 0; k < nblocks; ++k) {
            const int o = (int)(k & 1);
            const uint64_t r0 = (uint64_t)k * rb, rbk = std::min(rb, rows - r0);
            if (per_block && has_entries) {
                if (h_range[2 * k] <= h_range[2 * k + 1]) need_hi = std::max<uint64_t>(need_hi, (uint64_t)h_range[2 * k + 1] + 1);
            } else {
                need_hi = win_hi;
            }
            const uint64_t need_chunks = std::min(chunk_end, (need_hi + cb - 1) / cb);
            while (next_chunk < need_chunks) BSM_TRY(enqueue_chunk(next_chunk++));
            if (k >= 2) BSM_CUDA(cudaStreamWaitEvent(s_mm, ev_out_free[o], 0));
            if (k >= 2 && by_masks) BSM_CUDA(cudaStreamWaitEvent(s_mm, ev_landed[k - 2], 0));   // masks_st / rp64 of block k-2 have left
            bsm_dense cv = cd;
            cv.data = (char *)cdev + r0 * ld * s;
            cv.rows = rbk;
            if (per_block) {
                bsm_csr view = *a;   // rows [r0, r0 + rbk) of A: the same arrays, row_ptr window shifted (entries stay absolute)
                view.row_ptr = a->row_ptr + r0;
                view.rows = rbk;
                view.nnz = q.row_index[r0 + rbk] - q.row_index[r0];
                view.row_offset = a->row_offset + r0;
                view.owns = false;
                view.part_rows = nullptr;
                view.carry_vals = nullptr;
                view.long_rows = nullptr;
                view.part_chunks = 0;
                view.carry_vals_bytes = view.long_rows_cap = 0;
                tl.begin(PH_SPMM, s_mm);
                BSM_TRY(spmm_dispatch(&view, &bd, &cv, &tn, s_mm));
                tl.end(s_mm);
            } else if (k == 0) {
                tl.begin(PH_SPMM, s_mm);
                BSM_TRY(spmm_dispatch(a, &bd, &cd, &tn, s_mm));
                tl.end(s_mm);
            }
            if (want_csr) {
                // result construction of the block: insert's zero-drop (sparse.rs:229) + finalise's row_index (206-219)
                tl.begin(PH_COMPACT, s_mm);
                BSM_CUDA(cudaMemsetAsync(counts[o], 0, (pad4(rb + 1) + 4) * 4, s_mm));
                BSM_TRY(launch_count_nonzero(q.dtype, cv.data, rbk, n, ld, counts[o], nullptr, s_mm, by_masks ? masks_st[o] : nullptr));
                BSM_TRY(exclusive_scan_u32(counts[o], counts[o], rbk + 1, s_mm));
                BSM_TRY(launch_scatter_nonzero64(q.dtype, cv.data, rbk, n, ld, counts[o], vals_st[o], cols_st[o], s_mm));
                BSM_TRY(launch_row_index_piece(counts[o], rbk, tot, k, rp64[o], h_tot, s_mm));   // also stores the running total to the host
                tl.end(s_mm);
            } else {
                tl.begin(PH_C_TRANSPOSE, s_mm);
                BSM_TRY(launch_transpose_rm2cm(q.dtype, cv.data, stage_out[o], rbk, n, ld, s_mm));
                tl.end(s_mm);
            }
            BSM_CUDA(cudaEventRecord(ev_c[o], s_mm));
            if (by_masks) {   // the small pieces of the block leave at once, on their own stream: the host threads can start on its columns
                BSM_CUDA(cudaStreamWaitEvent(s_meta, ev_c[o], 0));
                BSM_CUDA(cudaMemcpyAsync(h_masks + r0 * mwords, masks_st[o], rbk * mwords * 8, cudaMemcpyDeviceToHost, s_meta));
                BSM_CUDA(cudaMemcpyAsync(q.out_row_index + r0, rp64[o], rbk * 8, cudaMemcpyDeviceToHost, s_meta));
                BSM_CUDA(cudaEventRecord(ev_landed[k], s_meta));
            }
            if (k >= 1) BSM_TRY(drain(k - 1));   // one block behind the compute
        }
        BSM_TRY(drain(nblocks - 1));
        {
            PhaseScope w(PH_WAIT);
            BSM_CUDA(cudaStreamSynchronize(s_out));
            BSM_CUDA(cudaStreamSynchronize(s_meta));
            BSM_CUDA(cudaStreamSynchronize(s_mm));
            BSM_CUDA(cudaStreamSynchronize(s_in));
        }
        if (want_csr) {
            q.out_row_index[rows] = h_tot[nblocks];   // finalise(): the tail of row_index is nnz
            *q.out_nnz = h_tot[nblocks];
        }
        {
            PhaseScope w(PH_WAIT);
            expand.finish();
        }
        return BSM_OK;
    }();
    expand.finish();   // (error paths: nothing may still be writing into the caller's arrays)
    for (cudaEvent_t e : ev_landed)
        if (e) cudaEventDestroy(e);
    if (st != BSM_OK && streams_touched) cudaDeviceSynchronize();   // nothing may still be using the buffers freed below
    tl.finish();
    for (int i = 0; i < 2; ++i) {
        tmp_free(stage_in[i]);
        tmp_free(stage_out[i]);
        tmp_free(vals_st[i]);
        tmp_free(cols_st[i]);
        tmp_free(masks_st[i]);
        tmp_free(rp64[i]);
        tmp_free(counts[i]);
    }
    tmp_free(bwin);
    tmp_free(cdev);
    tmp_free(tot);
    tmp_free(blk_range);
    bsm_csr_free(a);
    return st;
}

// The allocating form of the literal call: result arrays sized for the worst case (every output non-zero; untouched
// pages of a large malloc cost nothing), trimmed to the entries actually produced.
template <typename T>
int mul_dense_host_alloc(HostProduct<T> q, T **out_v, uint64_t **out_col_index, uint64_t **out_row_index)
{
    if (!q.out_nnz || !out_v || !out_col_index || !out_row_index) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense_host: null out");
    if (q.cols != q.rhs_rows) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "mul_dense: A.cols != rhs.rows (MatErr::IncorrectDimensions)");
    const uint64_t worst = q.rows * q.rhs_cols;
    T *hv = (T *)malloc(std::max<uint64_t>(1, worst) * sizeof(T));
    uint64_t *hc = (uint64_t *)malloc(std::max<uint64_t>(1, worst) * 8);
    uint64_t *hr = (uint64_t *)malloc((q.rows + 1) * 8);
    int st = (!hv || !hc || !hr) ? fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense_host: out of host memory") : BSM_OK;
    if (st == BSM_OK) {
        q.out_col_ptrs = nullptr;
        q.capacity = worst;
        q.out_v = hv;
        q.out_col_index = hc;
        q.out_row_index = hr;
        st = host_product<T>(q);
    }
    if (st != BSM_OK) {
        free(hv);
        free(hc);
        free(hr);
        return st;
    }
    const uint64_t nnz = std::max<uint64_t>(1, *q.out_nnz);
    if (nnz < worst) {   // shrinking never moves a large block; keep the original if realloc declines
        if (void *p = realloc(hv, nnz * sizeof(T))) hv = (T *)p;
        if (void *p = realloc(hc, nnz * 8)) hc = (uint64_t *)p;
    }
    *out_v = hv;
    *out_col_index = hc;
    *out_row_index = hr;
    return BSM_OK;
}

}  // namespace

template <typename T>
static int mul_vector(const bsm_csr *a, int dtype, const T *rhs, uint64_t rhs_len, T *out, uint64_t out_len)
{
    BSM_TRY(ensure_init());
    if (!a) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_vector: null handle");
    if (a->dtype != dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "mul_vector: dtype mismatch");
    // src/sparse.rs:469-471
    if (a->cols != rhs_len || a->rows != out_len)
        return fail(BSM_ERR_INCORRECT_DIMENSIONS, "mul_vector: dims (MatErr::IncorrectDimensions)");
    bsm_dense *x = nullptr, *y = nullptr;
    PoolScope pool;
    int st = [&]() -> int {
        BSM_TRY(dense_alloc(dtype, rhs_len, 1, &x));
        BSM_TRY(dense_alloc(dtype, out_len, 1, &y));
        if (rhs_len) BSM_CUDA(cudaMemcpyAsync(x->data, rhs, rhs_len * sizeof(T), cudaMemcpyHostToDevice, rt().stream));
        BSM_TRY(bsm_spmm(a, x, y, BSM_ALGO_AUTO));
        if (out_len) BSM_CUDA(cudaMemcpyAsync(out, y->data, out_len * sizeof(T), cudaMemcpyDeviceToHost, rt().stream));
        BSM_CUDA(cudaStreamSynchronize(rt().stream));
        return BSM_OK;
    }();
    if (x) bsm_dense_free(x);
    if (y) bsm_dense_free(y);
    return st;
}

}  // namespace bsm

using namespace bsm;

extern "C" {

int bsm_dense_residual_norm(const bsm_dense *ax, const bsm_dense *b, double *resid_fro, double *b_fro)
{
    BSM_TRY(ensure_init());
    if (!ax || !b || !resid_fro || !b_fro) return fail(BSM_ERR_INVALID_ARGUMENT, "residual_norm: null argument");
    if (ax->rows != b->rows || ax->cols != b->cols) return fail(BSM_ERR_INCORRECT_DIMENSIONS, "residual_norm: shapes differ");
    if (ax->dtype != b->dtype) return fail(BSM_ERR_DTYPE_MISMATCH, "residual_norm: dtype mismatch");
    const int nd = residual_norm_scratch_doubles();
    double *d = nullptr;
    BSM_TRY(tmp_alloc((void **)&d, nd * sizeof(double)));
    int st = launch_residual_norms(ax->dtype, ax->data, ax->ld, b->data, b->ld, ax->rows, ax->cols, d, rt().stream);
    double h[2] = {0.0, 0.0};
    if (st == BSM_OK) {
        cudaError_t e = cudaMemcpyAsync(h, d + nd - 2, 16, cudaMemcpyDeviceToHost, rt().stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(rt().stream);
        if (e != cudaSuccess) st = fail(BSM_ERR_CUDA, std::string("residual_norm: ") + cudaGetErrorString(e));
    }
    tmp_free(d);
    if (st != BSM_OK) return st;
    *resid_fro = std::sqrt(h[0]);
    *b_fro = std::sqrt(h[1]);
    return BSM_OK;
}

int bsm_dense_to_csr(const bsm_dense *d, bsm_csr **out) { return dense_to_csr_impl(d, out); }

#define BSM_HOST_PRODUCT(T, DT)                                                                                              \
    HostProduct<T> q{};                                                                                                      \
    q.dtype = DT;                                                                                                            \
    q.rows = rows;                                                                                                           \
    q.cols = cols;                                                                                                           \
    q.nnz = nnz;                                                                                                             \
    q.v = v;                                                                                                                 \
    q.col_index = col_index;                                                                                                 \
    q.row_index = row_index;                                                                                                 \
    q.row_index_len = row_index_len;                                                                                         \
    q.rhs_rows = rhs_rows;                                                                                                   \
    q.rhs_cols = rhs_cols;                                                                                                   \
    q.rhs_col_ptrs = rhs_col_ptrs;                                                                                           \
    q.algo = algo;

int bsm_mul_dense_host_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v, const uint64_t *col_index,
                           const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                           const double *const *rhs_col_ptrs, int algo, uint64_t *out_nnz, double **out_v,
                           uint64_t **out_col_index, uint64_t **out_row_index)
{
    BSM_HOST_PRODUCT(double, BSM_F64)
    q.out_nnz = out_nnz;
    return mul_dense_host_alloc<double>(q, out_v, out_col_index, out_row_index);
}
int bsm_mul_dense_host_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v, const uint64_t *col_index,
                           const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                           const float *const *rhs_col_ptrs, int algo, uint64_t *out_nnz, float **out_v,
                           uint64_t **out_col_index, uint64_t **out_row_index)
{
    BSM_HOST_PRODUCT(float, BSM_F32)
    q.out_nnz = out_nnz;
    return mul_dense_host_alloc<float>(q, out_v, out_col_index, out_row_index);
}
int bsm_mul_dense_host_into_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v, const uint64_t *col_index,
                                const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                const double *const *rhs_col_ptrs, int algo, uint64_t capacity, double *out_v,
                                uint64_t *out_col_index, uint64_t *out_row_index, uint64_t *out_nnz)
{
    BSM_HOST_PRODUCT(double, BSM_F64)
    q.capacity = capacity;
    q.out_v = out_v;
    q.out_col_index = out_col_index;
    q.out_row_index = out_row_index;
    q.out_nnz = out_nnz;
    return host_product<double>(q);
}
int bsm_mul_dense_host_into_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v, const uint64_t *col_index,
                                const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                const float *const *rhs_col_ptrs, int algo, uint64_t capacity, float *out_v,
                                uint64_t *out_col_index, uint64_t *out_row_index, uint64_t *out_nnz)
{
    BSM_HOST_PRODUCT(float, BSM_F32)
    q.capacity = capacity;
    q.out_v = out_v;
    q.out_col_index = out_col_index;
    q.out_row_index = out_row_index;
    q.out_nnz = out_nnz;
    return host_product<float>(q);
}
int bsm_mul_dense_host_dense_f64(uint64_t rows, uint64_t cols, uint64_t nnz, const double *v, const uint64_t *col_index,
                                 const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                 const double *const *rhs_col_ptrs, double *const *out_col_ptrs, int algo)
{
    if (rhs_cols && !out_col_ptrs) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense_host_dense: null column pointers");
    BSM_HOST_PRODUCT(double, BSM_F64)
    static double *const no_cols[1] = {nullptr};
    q.out_col_ptrs = out_col_ptrs ? out_col_ptrs : no_cols;
    return host_product<double>(q);
}
int bsm_mul_dense_host_dense_f32(uint64_t rows, uint64_t cols, uint64_t nnz, const float *v, const uint64_t *col_index,
                                 const uint64_t *row_index, uint64_t row_index_len, uint64_t rhs_rows, uint64_t rhs_cols,
                                 const float *const *rhs_col_ptrs, float *const *out_col_ptrs, int algo)
{
    if (rhs_cols && !out_col_ptrs) return fail(BSM_ERR_INVALID_ARGUMENT, "mul_dense_host_dense: null column pointers");
    BSM_HOST_PRODUCT(float, BSM_F32)
    static float *const no_cols[1] = {nullptr};
    q.out_col_ptrs = out_col_ptrs ? out_col_ptrs : no_cols;
    return host_product<float>(q);
}
void bsm_host_free(void *p) { free(p); }

int bsm_mul_vector_f64(const bsm_csr *a, const double *rhs, uint64_t rhs_len, double *out, uint64_t out_len)
{
    return mul_vector<double>(a, BSM_F64, rhs, rhs_len, out, out_len);
}
int bsm_mul_vector_f32(const bsm_csr *a, const float *rhs, uint64_t rhs_len, float *out, uint64_t out_len)
{
    return mul_vector<float>(a, BSM_F32, rhs, rhs_len, out, out_len);
}

}  // extern "C"
