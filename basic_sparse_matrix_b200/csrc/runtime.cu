// runtime.cu — process state of libbsm_b200.so: device and stream, memory pools, error strings, launch counter,
// the cached streams of the host-to-host pipelines, per-phase wall timers.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>

#include "bsm_internal.h"

namespace bsm {

static Runtime g_rt;
static thread_local std::string g_err;
static thread_local bsm_launch_info g_info;
static std::atomic<uint64_t> g_launches{0};

Runtime &rt() { return g_rt; }
void set_error(const std::string &msg) { g_err = msg; }
int fail(int status, const std::string &msg)
{
    g_err = msg;
    return status;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int ensure_init()
{
    if (g_rt.device >= 0) return BSM_OK;
    return bsm_init(0);
}

bsm_launch_info &launch_info() { return g_info; }

static thread_local int g_pool_depth = 0;
PoolScope::PoolScope() { ++g_pool_depth; }
PoolScope::~PoolScope() { --g_pool_depth; }

int tmp_alloc(void **p, size_t bytes)
{
    BSM_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, g_rt.stream));
    return BSM_OK;
}
void tmp_free(void *p)
{
    if (p) cudaFreeAsync(p, g_rt.stream);
}
int dev_alloc(void **p, size_t bytes, bool *pooled)
{
    *pooled = g_pool_depth > 0;
    if (*pooled) return tmp_alloc(p, bytes);
    BSM_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    return BSM_OK;
}
void dev_free(void *p, bool pooled)
{
    if (!p) return;
    if (pooled)
        cudaFreeAsync(p, g_rt.stream);
    else
        cudaFree(p);
}

// The three streams (host->device, multiply, device->host) and the events of the host-to-host pipelines live as long as
// the process: creating them costs tens of microseconds each and a call may be as short as a millisecond.
static PipelineStreams g_pipe;
static int g_pipe_device = -1;
int pipeline_streams(PipelineStreams **out)
{
    BSM_TRY(ensure_init());
    if (g_pipe_device != g_rt.device) {
        if (g_pipe_device >= 0) {
            cudaStreamDestroy(g_pipe.in);
            cudaStreamDestroy(g_pipe.mm);
            cudaStreamDestroy(g_pipe.out);
            cudaStreamDestroy(g_pipe.meta);
            for (cudaEvent_t &e : g_pipe.ev)
                if (e) cudaEventDestroy(e);
            g_pipe = PipelineStreams();
            g_pipe_device = -1;
        }
        BSM_CUDA(cudaStreamCreateWithFlags(&g_pipe.in, cudaStreamNonBlocking));
        BSM_CUDA(cudaStreamCreateWithFlags(&g_pipe.mm, cudaStreamNonBlocking));
        BSM_CUDA(cudaStreamCreateWithFlags(&g_pipe.out, cudaStreamNonBlocking));
        BSM_CUDA(cudaStreamCreateWithFlags(&g_pipe.meta, cudaStreamNonBlocking));
        for (cudaEvent_t &e : g_pipe.ev) BSM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g_pipe_device = g_rt.device;
    }
    *out = &g_pipe;
    return BSM_OK;
}

static void *g_pin = nullptr;
static size_t g_pin_bytes = 0;
int pinned_scratch(void **p, size_t bytes)
{
    if (g_pin_bytes < bytes) {
        if (g_pin) cudaFreeHost(g_pin);
        g_pin = nullptr;
        g_pin_bytes = 0;
        const size_t want = std::max<size_t>(bytes, 64 << 10);
        BSM_CUDA(cudaHostAlloc(&g_pin, want, cudaHostAllocMapped | cudaHostAllocPortable));   // mapped: kernels can store results into it
        g_pin_bytes = want;
    }
    *p = g_pin;
    return BSM_OK;
}

// ---- per-phase wall timers of the host-to-host calls -----------------------------------------------------------
static int g_phase_on = -1;   // -1 = read BSM_PHASE_TIMERS on first use
static double g_phase[PH_COUNT] = {};
bool phase_timers_on()
{
    if (g_phase_on < 0) {
        const char *e = getenv("BSM_PHASE_TIMERS");
        g_phase_on = (e && *e && *e != '0') ? 1 : 0;
    }
    return g_phase_on == 1;
}
void phase_add(int phase, double seconds)
{
    if (phase >= 0 && phase < PH_COUNT) g_phase[phase] += seconds;
}
double wall_seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace bsm

using namespace bsm;

extern "C" {

int bsm_abi_version(void) { return BSM_ABI_VERSION; }

int bsm_device_count(int *count)
{
    if (!count) return fail(BSM_ERR_INVALID_ARGUMENT, "device_count: null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(BSM_ERR_NO_DEVICE, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    }
    *count = n;
    return BSM_OK;
}

int bsm_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(BSM_ERR_NO_DEVICE, std::string("no usable CUDA device (this library has no CPU fallback): ") +
                                           (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(BSM_ERR_INVALID_ARGUMENT, "bsm_init: device index out of range");
    BSM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    BSM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (g_rt.own_stream && g_rt.device != device) {
        cudaStreamDestroy(g_rt.own_stream);
        g_rt.own_stream = nullptr;
    }
    if (!g_rt.own_stream) BSM_CUDA(cudaStreamCreateWithFlags(&g_rt.own_stream, cudaStreamNonBlocking));
    g_rt.stream = g_rt.own_stream;
    g_rt.device = device;
    g_rt.sm_count = prop.multiProcessorCount;
    g_rt.l2_bytes = (size_t)prop.l2CacheSize;
    g_rt.hbm_bytes = prop.totalGlobalMem;
    g_rt.cc_major = prop.major;
    g_rt.cc_minor = prop.minor;
    g_rt.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
        unsigned long long keep = ~0ull;   // never trim the stream-ordered pool behind our back
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return BSM_OK;
}

int bsm_set_stream(void *cuda_stream)
{
    BSM_TRY(ensure_init());
    g_rt.stream = cuda_stream ? (cudaStream_t)cuda_stream : g_rt.own_stream;
    return BSM_OK;
}

int bsm_sync(void)
{
    BSM_TRY(ensure_init());
    BSM_CUDA(cudaStreamSynchronize(g_rt.stream));
    return BSM_OK;
}

const char *bsm_last_error_string(void) { return g_err.c_str(); }

const char *bsm_status_string(int status)
{
    switch (status) {
        case BSM_OK: return "ok";
        case BSM_ERR_INCORRECT_DIMENSIONS: return "IncorrectDimensions";
        case BSM_ERR_NOT_FINALISED: return "MatrixNotFinalised";
        case BSM_ERR_OUT_OF_BOUNDS: return "OutOfBounds";
        case BSM_ERR_INDEX_OVERFLOW: return "IndexOverflow";
        case BSM_ERR_INVALID_ARGUMENT: return "InvalidArgument";
        case BSM_ERR_DTYPE_MISMATCH: return "DtypeMismatch";
        case BSM_ERR_CUDA: return "CudaError";
        case BSM_ERR_NCCL: return "NcclError";
        case BSM_ERR_NO_DEVICE: return "NoDevice";
        case BSM_ERR_NOT_SUPPORTED: return "NotSupported";
    }
    return "unknown";
}

int bsm_device_info(int *sm_count, size_t *l2_bytes, size_t *hbm_bytes, int *cc_major, int *cc_minor)
{
    BSM_TRY(ensure_init());
    if (sm_count) *sm_count = g_rt.sm_count;
    if (l2_bytes) *l2_bytes = g_rt.l2_bytes;
    if (hbm_bytes) *hbm_bytes = g_rt.hbm_bytes;
    if (cc_major) *cc_major = g_rt.cc_major;
    if (cc_minor) *cc_minor = g_rt.cc_minor;
    return BSM_OK;
}

uint64_t bsm_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int bsm_l2_flush(void)
{
    BSM_TRY(ensure_init());
    const size_t want = std::max<size_t>(g_rt.l2_bytes * 2, (size_t)256 << 20);
    if (g_rt.flush_bytes < want) {
        if (g_rt.flush_buf) cudaFree(g_rt.flush_buf);
        g_rt.flush_buf = nullptr;
        g_rt.flush_bytes = 0;
        BSM_CUDA(cudaMalloc(&g_rt.flush_buf, want));
        g_rt.flush_bytes = want;
    }
    BSM_CUDA(cudaMemsetAsync(g_rt.flush_buf, 0, g_rt.flush_bytes, g_rt.stream));
    return BSM_OK;
}

int bsm_phase_timers_enable(int on)
{
    g_phase_on = on ? 1 : 0;
    return BSM_OK;
}

int bsm_phase_timers_read(double *seconds, int count, int reset)
{
    if (!seconds || count < 0) return fail(BSM_ERR_INVALID_ARGUMENT, "phase_timers_read: bad arguments");
    for (int i = 0; i < count; ++i) seconds[i] = i < PH_COUNT ? g_phase[i] : 0.0;
    if (reset)
        for (double &x : g_phase) x = 0.0;
    return BSM_OK;
}

const char *bsm_phase_name(int phase)
{
    static const char *names[PH_COUNT] = {"a_upload_host", "a_stats_and_sync_host", "b_h2d_device", "b_transpose_device", "spmm_device", "compact_device",
                                          "c_transpose_device", "d2h_device", "wait_host", "total_host"};
    return phase >= 0 && phase < PH_COUNT ? names[phase] : "";
}

}  // extern "C"
