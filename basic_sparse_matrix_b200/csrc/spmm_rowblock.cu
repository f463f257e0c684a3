// spmm_rowblock.cu — vector-CSR variant for band-like matrices: blocks of consecutive rows whose stored
// columns are RUNS OF CONSECUTIVE INDICES (a band, dense diagonal blocks) share their B-row loads (sm_100a).
//
// Same contraction as Csr::mul_dense (/root/reference/src/sparse.rs:431-444). The plain vector kernel gathers
// one B row per stored entry: 65 gathers per output row of the half-bandwidth-32 matrix of BASELINE config 5,
// all of them L1 hits, and the L1 data pipe is the limit (0.28 of the HBM roofline). Here a lane group owns RB
// consecutive rows and walks the UNION of their column ranges once: B row j is loaded once and accumulated into
// every row of the block that stores column j (entry index = row start + j - first column, no col_idx read).
// Adjacent rows of a band share all but one column, so RB = 8 rows need 72 loads instead of 520.
//   * every row still sums its entries in stored order (ascending j IS the stored order of a run of consecutive
//     columns) with separately rounded multiply and add -> bit-identical to the reference for any values;
//   * rowblock_probe_kernel decides per matrix (cached in the handle) whether every row is such a run and how
//     much the blocks share; anything else stays on spmm_rows_kernel.
#include <algorithm>

#include "bsm_common.cuh"
#include "kernels.h"

namespace bsm {

// One thread per block of RB rows: are all rows runs of consecutive columns? How long is the union range?
// out[0] += union lengths (what the kernel will load), out[1] = number of rows that are not runs.
__global__ void rowblock_probe_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col_idx, uint32_t rows,
                                      uint32_t rb, unsigned long long *__restrict__ out)
{
    const uint64_t blk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t r0 = blk * rb;
    if (r0 >= rows) return;
    const uint32_t r1 = (uint32_t)min((uint64_t)rows, r0 + rb);
    uint32_t lo = 0xFFFFFFFFu, hi = 0, bad = 0;
    for (uint32_t r = (uint32_t)r0; r < r1; ++r) {
        const uint32_t s = row_ptr[r], e = row_ptr[r + 1];
        if (s == e) continue;
        const uint32_t first = col_idx[s];
        for (uint32_t k = s + 1; k < e; ++k) bad |= (col_idx[k] != first + (k - s));
        lo = min(lo, first);
        hi = max(hi, first + (e - s));
    }
    if (bad) atomicAdd(out + 1, 1ull);
    if (hi > lo) atomicAdd(out, (unsigned long long)(hi - lo));
}

// Shared memory: one stage of `cap` values per warp (the values of the warp's RPP * RB consecutive rows are one
// contiguous piece of the value array: a single TMA bulk copy, cp.async.bulk -> UBLKCP, completing on the warp's
// mbarrier) + the mbarriers. The value reads are then warp-broadcast LDS instead of scattered global loads.
// FUSED: the opt-in BSM_TUNE_FUSED arithmetic (one FMA per product; tolerance-level agreement with the reference)
// NT: register tiles per lane (tile t holds the columns t*G*V + gl*V ...): with two tiles a lane's value read feeds twice as many
// products (the kernel is bound by the instructions it issues per product; full-width shapes only)
template <typename T, int V, int G, int RB, bool FULLN, bool FUSED, int NT = 1>
__global__ void __launch_bounds__(256, (RB >= 8 || NT > 1) ? 2 : 3) spmm_rowblock_kernel(const RowBlockParams p)
{
    static_assert(NT == 1 || FULLN, "several tiles per lane: full-width shapes only");
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int RPP = 32 / G;                  // lane groups (row blocks) per warp
    constexpr uint32_t WR = RPP * RB;            // rows per warp block
    const uint32_t lane = threadIdx.x & 31, gl = lane % G, grp = lane / G, warp = threadIdx.x >> 5;
    const uint32_t warps = blockDim.x >> 5;
    const uint64_t warps_total = (uint64_t)gridDim.x * warps;
    T *stage = reinterpret_cast<T *>(smem) + (size_t)warp * p.cap;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)warps * p.cap * sizeof(T)) + warp;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncwarp();
    const bool col_ok = FULLN || gl * V < p.n;
    const T *__restrict__ vals = static_cast<const T *>(p.vals);
    const char *__restrict__ b_bytes = reinterpret_cast<const char *>(static_cast<const T *>(p.B) + gl * V);
    char *__restrict__ c_bytes = reinterpret_cast<char *>(static_cast<T *>(p.C) + gl * V);
    const uint32_t ldb_bytes = p.ldb * (uint32_t)sizeof(T), ldc_bytes = p.ldc * (uint32_t)sizeof(T);
    const bool streaming = (p.flags & BSM_TUNE_C_STREAMING) != 0;
    const uint64_t policy = (p.flags & BSM_TUNE_A_EVICT_FIRST) ? l2_policy_evict_first() : l2_policy_evict_normal();
    const unsigned long long negzero2 = packed_negzero(p.flags >> 31);   // (bit 31 of the flags is never set: the host masks it)
    uint32_t phase = 0;

    // neighbouring warps take neighbouring warp blocks: their B ranges overlap and meet in L1
    for (uint64_t wb = (uint64_t)blockIdx.x * warps + warp; wb * WR < p.rows; wb += warps_total) {
        const uint32_t wrow0 = (uint32_t)(wb * WR);
        const uint32_t wrows = min(WR, p.rows - wrow0);
        const uint32_t ws = __ldg(p.row_ptr + wrow0), we = __ldg(p.row_ptr + wrow0 + wrows);
        const uint32_t wbase = ws & ~3u;                       // 16-byte aligned start
        const uint32_t cnt = (we - wbase + 3u) & ~3u;          // entries, multiple of 4
        __syncwarp();   // every lane is done reading the stage
        if (lane == 0 && we > ws) {
            if (cnt > p.cap) __trap();   // the host sizes cap from the longest row
            mbar_arrive_expect_tx(bar, cnt * (uint32_t)sizeof(T));
            bulk_g2s(stage, vals + wbase, cnt * (uint32_t)sizeof(T), bar, policy);
        }
        // this group's rows, while the copy is in flight
        const uint32_t row0 = wrow0 + grp * RB;
        uint32_t base[RB], first[RB], len[RB];
        uint32_t jlo = 0xFFFFFFFFu, jhi = 0;        // union of the rows' column ranges
        uint32_t ja = 0, jb = 0xFFFFFFFFu;          // their intersection [ja, jb): every row stores these columns
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const uint32_t row = row0 + r;
            base[r] = first[r] = len[r] = 0;
            if (row < p.rows) {
                base[r] = __ldg(p.row_ptr + row);
                len[r] = __ldg(p.row_ptr + row + 1) - base[r];
            }
            if (len[r]) {
                first[r] = __ldg(p.col_idx + base[r]);
                jlo = min(jlo, first[r]);
                jhi = max(jhi, first[r] + len[r]);
                ja = max(ja, first[r]);
                jb = min(jb, first[r] + len[r]);
            } else {
                jb = 0;   // an empty (or missing) row: no common columns
            }
        }
        if (jb <= ja) ja = jb = jhi;   // no intersection: the predicated loop walks the whole union
        Lane<T, V> acc[RB][NT];
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[r][t].zero();   // T::default()  sparse.rs:434
        if (we > ws) {
            mbar_wait(bar, phase);   // the values have landed
            phase ^= 1u;
        }
        const T *va = stage - wbase;   // entry k of the matrix at va[k]
        // columns only some of the rows store (head and tail of the union): predicated
        auto ragged = [&](uint32_t j0, uint32_t j1) {
            const char *brow = b_bytes + (size_t)j0 * ldb_bytes;
            for (uint32_t j = j0; j < j1; ++j, brow += ldb_bytes) {
                Lane<T, V> b[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    b[t].zero();
                    if (col_ok) b[t].template load<false>(reinterpret_cast<const T *>(brow) + t * G * V, 0ull);
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const uint32_t off = j - first[r];
                    if (off < len[r]) {   // row r stores column j, at entry base + off; ascending j = stored order
                        const T a = va[base[r] + off];
#pragma unroll
                        for (int t = 0; t < NT; ++t) axpy<FUSED, T, V>(a, b[t].x, acc[r][t].x, negzero2);
                    }
                }
            }
        };
        if (jhi > jlo) {
            ragged(jlo, ja);
            // columns every row of the block stores (all but RB-1 at each end of a band): no predicates, so the
            // loads of several columns are in flight together
            {
                const T *vrow[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) vrow[r] = va + base[r] - first[r];   // entry of column j at vrow[r][j]
                const char *brow = b_bytes + (size_t)ja * ldb_bytes;
#pragma unroll 4
                for (uint32_t j = ja; j < jb; ++j, brow += ldb_bytes) {
                    Lane<T, V> b[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        b[t].zero();
                        if (col_ok) b[t].template load<false>(reinterpret_cast<const T *>(brow) + t * G * V, 0ull);
                    }
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        const T a = vrow[r][j];
#pragma unroll
                        for (int t = 0; t < NT; ++t) axpy<FUSED, T, V>(a, b[t].x, acc[r][t].x, negzero2);
                    }
                }
            }
            ragged(jb, jhi);
        }
#pragma unroll
        for (int r = 0; r < RB; ++r)
            if (row0 + r < p.rows && col_ok) {
#pragma unroll
                for (int t = 0; t < NT; ++t) acc[r][t].store(reinterpret_cast<T *>(c_bytes + (size_t)(row0 + r) * ldc_bytes) + t * G * V, streaming);
            }
    }
}

int launch_rowblock_probe(const uint32_t *row_ptr, const uint32_t *col_idx, uint64_t rows, unsigned long long *out, cudaStream_t stream)
{
    const uint64_t blocks_of_rows = (rows + kRowBlockRows - 1) / kRowBlockRows;
    if (blocks_of_rows == 0) return BSM_OK;
    const uint32_t threads = 128;
    rowblock_probe_kernel<<<(uint32_t)((blocks_of_rows + threads - 1) / threads), threads, 0, stream>>>(row_ptr, col_idx, (uint32_t)rows,
                                                                                                         kRowBlockRows, out);
    BSM_CUDA(cudaGetLastError());
    count_launch();
    return BSM_OK;
}

// Instantiations: 128-bit lanes (f64 x 2, f32 x 4) and one-element lanes (operands that are not 16-byte granular),
// 4 / 8 / 16 / 32 lanes per row, blocks of 4 or 8 rows; the fused arithmetic for full-width 128-bit shapes only.
// Anything else (2-element f32 lanes, 1 or 2 lanes per row) reports "no kernel": BSM_ALGO_AUTO then runs the plain vector kernel.
// Blocks of 8 rows exist where they measured faster: a full warp per row with 128-bit lanes (r1_sweepag_band_*); else 4.
template <typename T, int V, int G> constexpr bool rowblock_has_rb8() { return G == 32 && V * sizeof(T) == 16; }
template <typename T, int V, int G> static const void *rowblock_ptr(bool fulln, int rb, bool fused)
{
    if constexpr (rowblock_has_rb8<T, V, G>()) {
        if (rb == 8) {
            if (fused) return fulln ? reinterpret_cast<const void *>(&spmm_rowblock_kernel<T, V, G, 8, true, true>) : nullptr;
            return fulln ? reinterpret_cast<const void *>(&spmm_rowblock_kernel<T, V, G, 8, true, false>)
                         : reinterpret_cast<const void *>(&spmm_rowblock_kernel<T, V, G, 8, false, false>);
        }
    }
    if constexpr (V * sizeof(T) == 16) {
        if (fused && fulln) return reinterpret_cast<const void *>(&spmm_rowblock_kernel<T, V, G, 4, true, true>);
    }
    if (fused) return nullptr;
    return fulln ? reinterpret_cast<const void *>(&spmm_rowblock_kernel<T, V, G, 4, true, false>)
                 : reinterpret_cast<const void *>(&spmm_rowblock_kernel<T, V, G, 4, false, false>);
}
// two register tiles per lane: full-width 128-bit shapes, blocks of 4 rows (four tiles spill at 128 registers and measured slower:
// band x64 f64 0.87 -> 1.18 ms, x128 f32 0.87 -> 1.08 ms)
template <typename T, int V, int G, int NT> static const void *rowblock_tiles_ptr(bool fused)
{
    return fused ? reinterpret_cast<const void *>(&spmm_rowblock_kernel<T, V, G, 4, true, true, NT>)
                 : reinterpret_cast<const void *>(&spmm_rowblock_kernel<T, V, G, 4, true, false, NT>);
}
template <typename T, int V> static const void *rowblock_select_tiles(int G, int NT, bool fused)
{
    if constexpr (V * sizeof(T) == 16) {
        if (NT == 2) switch (G) {
                case 16: return rowblock_tiles_ptr<T, V, 16, 2>(fused);
                case 8: return rowblock_tiles_ptr<T, V, 8, 2>(fused);
                case 4: return rowblock_tiles_ptr<T, V, 4, 2>(fused);
            }
    }
    return nullptr;
}
template <typename T, int V> static const void *rowblock_select_g(int G, bool fulln, int rb, bool fused)
{
    switch (G) {
        case 32: return rowblock_ptr<T, V, 32>(fulln, rb, fused);
        case 16: return rowblock_ptr<T, V, 16>(fulln, rb, fused);
        case 8: return rowblock_ptr<T, V, 8>(fulln, rb, fused);
        case 4: return rowblock_ptr<T, V, 4>(fulln, rb, fused);
    }
    return nullptr;
}

int launch_spmm_rowblock(int dtype, Shape sh, const RowBlockParams &p_in, int rb, uint64_t max_row_nnz, int sm_count, size_t smem_max, cudaStream_t stream,
                         int *grid_out, int *block_out, int *smem_out, int *rb_out)
{
    if (sh.NT != 1 && sh.NT != 2) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_rowblock: one or two register tiles per lane");
    const bool fulln = p_in.n == (uint32_t)(sh.V * sh.G * sh.NT);
    if (sh.NT > 1 && !fulln) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_rowblock: several register tiles per lane exist for full-width shapes only");
    const void *k = nullptr;
    const bool fused = (p_in.flags & BSM_TUNE_FUSED) != 0;
    if (rb != 4 && rb != 8) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_rowblock: 4 or 8 rows per block");
    if (rb == 8 && !(sh.G == 32 && sh.V * dtype_size(dtype) == 16)) rb = 4;   // blocks of 8 rows: full-warp 128-bit shapes only
    if (sh.NT > 1) rb = 4;
    if (rb_out) *rb_out = rb;
    if (sh.NT > 1) {
        if (dtype == BSM_F64 && sh.V == 2) k = rowblock_select_tiles<double, 2>(sh.G, sh.NT, fused);
        if (dtype == BSM_F32 && sh.V == 4) k = rowblock_select_tiles<float, 4>(sh.G, sh.NT, fused);
    } else if (dtype == BSM_F64) {
        if (sh.V == 1) k = rowblock_select_g<double, 1>(sh.G, fulln, rb, fused);
        if (sh.V == 2) k = rowblock_select_g<double, 2>(sh.G, fulln, rb, fused);
    } else {
        if (sh.V == 1) k = rowblock_select_g<float, 1>(sh.G, fulln, rb, fused);
        if (sh.V == 4) k = rowblock_select_g<float, 4>(sh.G, fulln, rb, fused);
    }
    if (!k) return fail(BSM_ERR_NOT_SUPPORTED, fused ? "spmm_rowblock: BSM_TUNE_FUSED exists for full-width 128-bit lane shapes only" : "spmm_rowblock: no kernel for this lane shape");
    RowBlockParams pc = p_in;
    const uint64_t wr = (uint64_t)(32 / sh.G) * (uint64_t)rb;                  // rows per warp block
    const uint64_t cap = ((wr * max_row_nnz + 3 + 3) & ~3ull) + 4;             // + aligned start, rounded size
    int block = 256;
    auto smem_of = [&](int threads) { return (size_t)(threads / 32) * (cap * dtype_size(dtype) + 8); };
    while (block > 32 && smem_of(block) > smem_max) block /= 2;
    const size_t smem = smem_of(block);
    if (smem > smem_max || cap >= 0xFFFFFFF0ull) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_rowblock: the rows of one warp block do not fit shared memory");
    pc.cap = (uint32_t)cap;
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    BSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, block, smem));
    if (occ < 1) return fail(BSM_ERR_CUDA, "spmm_rowblock: kernel does not fit on an SM");
    const uint64_t warp_blocks = ((uint64_t)pc.rows + wr - 1) / wr;
    const uint64_t want = (warp_blocks + (block / 32) - 1) / (block / 32);
    const int grid = (int)std::min<uint64_t>(want, (uint64_t)sm_count * occ);
    if (grid_out) *grid_out = grid;
    if (block_out) *block_out = block;
    if (smem_out) *smem_out = (int)smem;
    if (grid == 0) return BSM_OK;
    void *args[] = {&pc};
    BSM_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(block), args, smem, stream));
    count_launch();
    return BSM_OK;
}

}  // namespace bsm
