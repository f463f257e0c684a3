// bsm_internal.h — what the host-side translation units of libbsm_b200.so share (internal interface):
//   runtime.cu   process state (device, stream, memory pools, error strings, launch counter, phase timers)
//   handles.cu   bsm_csr / bsm_dense handles: allocation, upload / download with format conversion, statistics
//   planner.cu   pure host arithmetic: lane shapes, the vector kernel's launch plan, row partitioning
//   dispatch.cu  which kernel runs a product and with what geometry (bsm_spmm*, bsm_spmm_scatter)
//   pipeline.cu  host-to-host calls: the literal Csr::mul_dense (-> zero-dropped Csr), the dense-result twin, mul_vector
//   solve.cu     forward / backward substitution of lib.rs:28-65 on the device
//   gen_api.cu   synthetic workloads generated in HBM
#pragma once
#include <string>

#include "bsm_common.cuh"
#include "kernels.h"

namespace bsm {

// ---- runtime.cu --------------------------------------------------------------------------------------------
bsm_launch_info &launch_info();   // what the last bsm_spmm* call on this thread launched (thread-local)

// Two kinds of device memory:
//   * long-lived handles created by the public upload / alloc / generator calls: cudaMalloc;
//   * temporaries, and the handles the host-to-host convenience calls create and destroy inside one
//     call (PoolScope): the stream-ordered pool of the device, kept warm (release threshold = max), so
//     that a small multiplication does not pay a dozen cudaMalloc / cudaFree round trips.
struct PoolScope {
    PoolScope();
    ~PoolScope();
};
int tmp_alloc(void **p, size_t bytes);   // stream-ordered, on the library stream
void tmp_free(void *p);
int dev_alloc(void **p, size_t bytes, bool *pooled);
void dev_free(void *p, bool pooled);

// Streams and events of the host-to-host pipelines, created once per process (not per call).
struct PipelineStreams {
    cudaStream_t in = nullptr, mm = nullptr, out = nullptr, meta = nullptr;   // meta: the small pieces of a result block (row masks, row_index)
    cudaEvent_t ev[16] = {};
};
int pipeline_streams(PipelineStreams **out);
// a small pinned, device-mapped host buffer owned by the runtime (grown on demand, never shrunk): readbacks without a cudaHostAlloc
// per call, and scalars a kernel stores straight into host memory (unified addressing: the host pointer is valid on the device)
int pinned_scratch(void **p, size_t bytes);

// Per-phase wall-clock timers of the host-to-host calls (BSM_PHASE_TIMERS=1, or bsm_phase_timers_enable):
// accumulated per process, read and reset by bsm_phase_timers_read.
enum Phase { PH_A_UPLOAD = 0, PH_A_STATS, PH_B_H2D, PH_B_TRANSPOSE, PH_SPMM, PH_COMPACT, PH_C_TRANSPOSE, PH_D2H, PH_WAIT, PH_TOTAL, PH_COUNT };
bool phase_timers_on();
void phase_add(int phase, double seconds);
double wall_seconds();
struct PhaseScope {   // wall time between construction and destruction, when the timers are on
    int phase;
    double t0;
    explicit PhaseScope(int ph) : phase(ph), t0(phase_timers_on() ? wall_seconds() : 0.0) {}
    ~PhaseScope()
    {
        if (phase_timers_on()) phase_add(phase, wall_seconds() - t0);
    }
};

// ---- handles.cu --------------------------------------------------------------------------------------------
inline uint64_t pad4(uint64_t n) { return (n + 3) / 4 * 4; }
uint64_t default_ld(uint64_t cols, int dtype);
int alloc_csr(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, bsm_csr **out);
int dense_alloc(int dtype, uint64_t rows, uint64_t cols, bsm_dense **out);
// max row length, column range, row_ptr sanity, line length of a stencil-like matrix: ONE kernel, one readback
int compute_stats(bsm_csr *a, bool check_cols);
template <typename T>
int csr_upload(int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const T *v, const uint64_t *col_index, const uint64_t *row_index,
               uint64_t row_index_len, bsm_csr **out);
template <typename T> int csr_download(const bsm_csr *a, int dtype, T *v, uint64_t *col_index, uint64_t *row_index);
template <typename T> int dense_upload(int dtype, uint64_t rows, uint64_t cols, const T *const *col_ptrs, bsm_dense **out);

// ---- planner.cu (pure host arithmetic) -----------------------------------------------------------------------
// lane shape for `n` columns starting at byte-aligned pointers
Shape pick_shape(uint32_t n, uint64_t ldb, uint64_t ldc, uint64_t col0, const void *b, const void *c, size_t s, bool prefer_wide,
                 uint64_t extra_ld = 0);
// Columns one pass can take starting at col0 when `want` remain (<= the tile): the lane shape holds at most
// G * V * NT columns, and alignment can force vectors narrower than 16 bytes — 129 f32 columns are 129 one-element
// lanes, more than the 128 that four register tiles hold. Then the pass is cut to a width the widest vectors divide
// (128 of the 129; the last column goes to the next pass), or failing that to what the narrow vectors hold.
template <typename ShapeOf> uint32_t fit_pass_width(uint32_t want, int vmax, ShapeOf &&shape_of)
{
    auto holds = [](const Shape &x) { return (uint32_t)(x.G * x.V * x.NT); };
    const Shape sh = shape_of(want);
    if (holds(sh) >= want) return want;
    const uint32_t even = want / (uint32_t)vmax * (uint32_t)vmax;
    if (even >= (uint32_t)vmax && even < want && holds(shape_of(even)) >= even) return even;
    return holds(sh);
}
// What the planner knows about the matrix and the device — no pointers, no CUDA calls, so the same code plans a
// launch for bsm_spmm and answers bsm_plan_vector (a dry run the CPU test-suite uses to pin the heuristics).
struct MatrixFacts {
    int dtype;
    uint64_t rows, nnz, max_row_nnz;
    uint32_t row_stride;   // line length of a stencil-like matrix (0 = none)
    double mean() const { return rows ? (double)nnz / (double)rows : 0.0; }
};
struct DeviceFacts {
    int sm_count;
    size_t smem_max;       // dynamic shared memory one CTA may use
};
struct PassAlign {         // what pick_shape needs to know about the operands of one column pass
    uint64_t ldb, ldc, col0;
    const void *b, *c;
};
struct VectorPlan {
    Shape sh;
    bool grouped = false;
    int flavour = 0, nw = 0;
    uint32_t R = 0, P = 0, stages = 0, cap = 0, num_super = 0;
    size_t smem = 0;
    int resident = 1;      // CTAs per SM the flavour targets
};
int plan_vector_pass(const MatrixFacts &m, const DeviceFacts &dev, const bsm_tuning &tn, uint32_t n, const PassAlign &al, bool scatter, bool multi,
                     VectorPlan *out);
// width of the column tile one pass covers at most (bsm_tuning.col_tile, clamped to what `tiles_max` register tiles hold)
uint32_t column_tile(const bsm_tuning &tn, uint32_t n_total, int vmax, uint32_t tiles_max);

// ---- dispatch.cu -------------------------------------------------------------------------------------------
// C = A * B on `stream` (the C-ABI entry points pass the library stream; the pipelines pass their own)
int spmm_dispatch(const bsm_csr *a, const bsm_dense *b, bsm_dense *c, const bsm_tuning *tuning, cudaStream_t stream);
// does every row of the matrix store a run of consecutive columns? (the row-block probe: one kernel per handle, cached)
int csr_rows_are_runs(const bsm_csr *a, cudaStream_t stream, bool *runs);
// which kernel family bsm_spmm would run for `requested` (bsm_algo; AUTO = the heuristics) on n_cols columns
int resolve_algo(const bsm_csr *a, uint64_t n_cols, int requested, int *algo, cudaStream_t stream);

// ---- pipeline.cu -------------------------------------------------------------------------------------------
int dense_to_csr_impl(const bsm_dense *d, bsm_csr **out);

}  // namespace bsm
