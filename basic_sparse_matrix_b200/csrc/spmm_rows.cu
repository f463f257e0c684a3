// spmm_rows.cu — host side of the vector-CSR kernel: shared-memory sizing, occupancy, launch.
// The kernel itself is in spmm_rows_kernel.cuh; its instantiations in spmm_rows_f64.cu / _f32.cu.
#include <algorithm>

#include "spmm_rows_kernel.cuh"

namespace bsm {

const void *row_kernel_select_f64(Shape sh, bool fulln, int flavour, bool multi);
const void *row_kernel_select_f32(Shape sh, bool fulln, int flavour, bool multi);

size_t row_kernel_smem_bytes(int dtype, const RowParams &p, int warps)
{
    const size_t ring = (size_t)row_layout(p.cap, p.R, (uint32_t)dtype_size(dtype)).stage_bytes * p.stages * warps;
    return ring + (size_t)warps * p.stages * 8;   // + one mbarrier per (warp, stage)
}

static const void *row_kernel_select(int dtype, Shape sh, uint32_t n, int flavour, bool multi)
{
    const bool fulln = n == (uint32_t)(sh.V * sh.G * sh.NT);
    return dtype == BSM_F64 ? row_kernel_select_f64(sh, fulln, flavour, multi) : row_kernel_select_f32(sh, fulln, flavour, multi);
}

int row_kernel_occupancy(int dtype, Shape sh, uint32_t n, int flavour, bool multi, int block, size_t smem, int *blocks_per_sm)
{
    const void *k = row_kernel_select(dtype, sh, n, flavour, multi);
    if (!k) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_rows: no kernel for this lane shape");
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // query under the largest carve-out (launch_spmm_rows narrows it to what the chosen residency needs)
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    BSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, block, smem));
    return BSM_OK;
}

int launch_spmm_rows(int dtype, Shape sh, const RowParams &p, int flavour, bool multi, int grid, int block, size_t smem, int ctas_per_sm,
                     cudaStream_t stream)
{
    const void *k = row_kernel_select(dtype, sh, p.n, flavour, multi);
    if (!k) return fail(BSM_ERR_NOT_SUPPORTED, "spmm_rows: no kernel for this lane shape");
    if (p.stages < 1 || p.stages > (uint32_t)kMaxStages) return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_rows: stages out of range");
    const bool flat = sh.G == 32 || sh.NT > 1 || flavour == 8;   // flat-stream shapes accept any P >= R, the row-by-row ones need whole slices
    if (p.R == 0 || p.R % 4 || p.P < p.R || (!flat && p.P % p.R))
        return fail(BSM_ERR_INVALID_ARGUMENT, "spmm_rows: inconsistent slice geometry");
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // smallest shared-memory carve-out that still holds the resident CTAs' rings: the rest of the
    // 228 KB stays L1, which is what the B-row gathers live on
    const size_t need = (size_t)std::max(1, ctas_per_sm) * (smem + 1024);
    const int pct = (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024));
    BSM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    RowParams pc = p;
    void *args[] = {&pc};
    BSM_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(block), args, smem, stream));
    count_launch();
    return BSM_OK;
}

}  // namespace bsm
