// bsm_common.cuh — shared device helpers and host-side handle definitions (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "bsm.h"   // include/bsm.h (-I../../include here; next to the sources in the Rust crate)

namespace bsm {

// ------------------------------------------------------------------------------------------
// host-side state (bsm_api.cu)
// ------------------------------------------------------------------------------------------
struct Runtime {
    int device = -1;
    cudaStream_t stream = nullptr;      // stream all work is enqueued on
    cudaStream_t own_stream = nullptr;  // created by bsm_init
    int sm_count = 0;
    size_t l2_bytes = 0;
    size_t hbm_bytes = 0;
    int cc_major = 0, cc_minor = 0;
    int max_smem_optin = 0;
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
};
Runtime &rt();
void set_error(const std::string &msg);
int fail(int status, const std::string &msg);
int ensure_init();
void count_launch(int n = 1);

#define BSM_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return ::bsm::fail(BSM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define BSM_TRY(expr)                \
    do {                             \
        int _s = (expr);             \
        if (_s != BSM_OK) return _s; \
    } while (0)

inline size_t dtype_size(int dtype) { return dtype == BSM_F64 ? 8 : 4; }
inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

}  // namespace bsm

// Device-resident CSR. All three arrays are padded so that reads up to the next multiple of
// 4 entries past the logical end stay inside the allocation (TMA bulk copies are issued in
// 16-byte units from 16-byte aligned addresses).
struct bsm_csr {
    int dtype = BSM_F64;
    uint64_t rows = 0, cols = 0, nnz = 0;
    void *vals = nullptr;          // T[nnz]
    uint32_t *col_idx = nullptr;   // u32[nnz]
    uint32_t *row_ptr = nullptr;   // u32[rows+1]
    bool owns = true;
    bool pooled = false;           // arrays come from the stream-ordered pool (freed with cudaFreeAsync)
    bool cache_pooled = false;     // likewise for the merge-path caches below
    uint64_t max_row_nnz = 0;      // csr_row_stats (dispatch heuristic)
    uint32_t col_min = 0, col_max = 0;   // smallest / largest stored column (valid when nnz > 0)
    uint64_t row_offset = 0;       // global index of local row 0 (row blocks of a partitioned matrix keep global columns)
    uint32_t row_stride = 0;       // dominant off-diagonal column stride of a stencil-like matrix (0 = none)
    // row-block probe (spmm_rowblock.cu), run on first use: 0 = not probed, 1 = every row is a run of consecutive
    // columns, 2 = not; rowblock_union = B rows the row-block kernel would load (vs nnz for the vector kernel)
    int rowblock_state = 0;
    uint64_t rowblock_union = 0;
    // band probe (solve.cu), run on first use by a substitution: 0 = not probed, 1 = probed. band_lower / band_upper: the matrix is a
    // proper lower / upper band factor of half-bandwidth band_hb (every row stores exactly the band's columns, diagonal last / first)
    int band_state = 0;
    bool band_lower = false, band_upper = false;
    uint32_t band_hb = 0;
    // merge-path partition cache (depends only on A and the item count)
    int part_items = 0;
    uint32_t part_chunks = 0;
    uint32_t *part_rows = nullptr;   // u32[part_chunks+1]: first row each chunk closes
    // carry-out scratch of the merge kernel, grown on demand
    void *carry_vals = nullptr;
    uint32_t *long_rows = nullptr;   // [long_cap][2] + counter at the end (merge fix-up of hub rows)
    size_t carry_vals_bytes = 0, long_rows_cap = 0;
};

// Device-resident dense matrix, ROW-major with leading dimension ld (elements).
struct bsm_dense {
    int dtype = BSM_F64;
    uint64_t rows = 0, cols = 0, ld = 0;
    void *data = nullptr;
    bool owns = true;
    bool pooled = false;
    bool ipc = false;              // mapping of another process's buffer (cudaIpcOpenMemHandle)
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace bsm {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier (shared::cta) ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make barrier initialisation visible to the async proxy (TMA unit) before first use
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- TMA bulk copy global -> shared (cp.async.bulk; SASS: UBLKCP) ----------------------------
// dst/src 16-byte aligned, bytes a non-zero multiple of 16; completion is signalled on `bar`
// as transaction bytes.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---- per-lane column vectors -----------------------------------------------------------------
template <typename T, int V> struct VecT;
template <> struct VecT<double, 1> { using type = double; };
template <> struct VecT<double, 2> { using type = double2; };
template <> struct VecT<float, 1> { using type = float; };
template <> struct VecT<float, 2> { using type = float2; };
template <> struct VecT<float, 4> { using type = float4; };

// read-only (non-coherent) global loads of one lane's vector with an L2 eviction-priority hint
__device__ __forceinline__ double ldg_hint(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double2 ldg_hint(const double2 *p, uint64_t pol)
{
    double2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ldg_hint(const float *p, uint64_t pol)
{
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ldg_hint(const float2 *p, uint64_t pol)
{
    float2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ldg_hint(const float4 *p, uint64_t pol)
{
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

template <typename T, int V> struct Lane {
    using VT = typename VecT<T, V>::type;
    T x[V];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int i = 0; i < V; ++i) x[i] = T(0);
    }
    // HINT: carry an L2 eviction-priority policy (createpolicy) on the read-only load
    template <bool HINT> __device__ __forceinline__ void load(const T *p, uint64_t pol)
    {
        VT v;
        if constexpr (HINT)
            v = ldg_hint(reinterpret_cast<const VT *>(p), pol);
        else
            v = __ldg(reinterpret_cast<const VT *>(p));
        *reinterpret_cast<VT *>(x) = v;
    }
    __device__ __forceinline__ void load_plain(const T *p) { *reinterpret_cast<VT *>(x) = *reinterpret_cast<const VT *>(p); }
    __device__ __forceinline__ void store(T *p, bool streaming) const
    {
        VT v = *reinterpret_cast<const VT *>(x);
        if (streaming)
            __stcs(reinterpret_cast<VT *>(p), v);
        else
            *reinterpret_cast<VT *>(p) = v;
    }
};

// value = value + (a*b): multiply and add rounded separately, exactly like the reference
// (src/sparse.rs:438-439; rustc never contracts). The _rn intrinsics are never fused by nvcc.
__device__ __forceinline__ double mul_add_unfused(double a, double b, double acc) { return __dadd_rn(acc, __dmul_rn(a, b)); }
__device__ __forceinline__ float mul_add_unfused(float a, float b, float acc) { return __fadd_rn(acc, __fmul_rn(a, b)); }
__device__ __forceinline__ double mul_add_fused(double a, double b, double acc) { return fma(a, b, acc); }
__device__ __forceinline__ float mul_add_fused(float a, float b, float acc) { return fmaf(a, b, acc); }

// acc += a * b over a lane's V elements, product and sum rounded separately (as mul_add_unfused). f32 with an even V
// uses Blackwell's packed f32x2 pipe at two instructions per two elements:
//     q   = fma.rn.f32x2(a, b, -0.0)   == rn(a*b) exactly (an FMA whose addend is -0.0 rounds the exact product once
//                                         and keeps its sign, also for zero products) -> FFMA2
//     acc = add.rn.f32x2(acc, q)        -> FADD2
// A packed MULTIPLY feeding the packed add is NOT usable: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2,
// a fused multiply-add with a single rounding. `negzero2` = two -0.0f, built from a value the compiler cannot see
// through (else it folds x + -0.0 and contracts again); tests/test_build_flags.py checks the shipped SASS
// (FFMA2 and FADD2 in equal numbers, no scalar FFMA), the bitwise parity tests check the arithmetic.
template <typename T, int V> __device__ __forceinline__ void axpy_unfused(T a, const T (&b)[V], T (&acc)[V], unsigned long long negzero2)
{
    if constexpr (sizeof(T) == 4 && V % 2 == 0) {
#pragma unroll
        for (int i = 0; i < V; i += 2) {
            unsigned long long aa, bb, cc, qq;
            asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
            asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b[i]), "f"(b[i + 1]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(acc[i]), "f"(acc[i + 1]));
            asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(qq) : "l"(aa), "l"(bb), "l"(negzero2));
            asm("add.rn.f32x2 %0, %1, %2;" : "=l"(cc) : "l"(cc), "l"(qq));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(cc));
        }
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = mul_add_unfused(a, b[i], acc[i]);
    }
}
// OPT-IN fused form (BSM_TUNE_FUSED): acc = fma(a, b, acc), one rounding per product-and-sum; f32 pairs on the packed pipe (FFMA2)
template <typename T, int V> __device__ __forceinline__ void axpy_fused(T a, const T (&b)[V], T (&acc)[V])
{
    if constexpr (sizeof(T) == 4 && V % 2 == 0) {
#pragma unroll
        for (int i = 0; i < V; i += 2) {
            unsigned long long aa, bb, cc;
            asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
            asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b[i]), "f"(b[i + 1]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(acc[i]), "f"(acc[i + 1]));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(aa), "l"(bb));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(cc));
        }
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = mul_add_fused(a, b[i], acc[i]);
    }
}
template <bool FUSED, typename T, int V>
__device__ __forceinline__ void axpy(T a, const T (&b)[V], T (&acc)[V], unsigned long long negzero2)
{
    if constexpr (FUSED)
        axpy_fused<T, V>(a, b, acc);
    else
        axpy_unfused<T, V>(a, b, acc, negzero2);
}
// two -0.0f; `runtime_zero` must be 0 at run time and opaque at compile time (a kernel parameter bit that is never set)
__device__ __forceinline__ unsigned long long packed_negzero(uint32_t runtime_zero)
{
    unsigned long long r;
    const uint32_t nz = 0x80000000u | runtime_zero;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "r"(nz));
    return r;
}

template <bool FUSED, typename T> __device__ __forceinline__ T mul_add(T a, T b, T acc)
{
    if constexpr (FUSED)
        return mul_add_fused(a, b, acc);
    else
        return mul_add_unfused(a, b, acc);
}

}  // namespace bsm
#endif  // __CUDACC__
