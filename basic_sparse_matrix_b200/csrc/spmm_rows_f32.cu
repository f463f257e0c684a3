// spmm_rows_f32.cu — f32 instantiations of the vector-CSR kernel (see spmm_rows_kernel.cuh).
#include "spmm_rows_inst.cuh"

namespace bsm {
const void *row_kernel_select_f32(Shape sh, bool fulln, int flavour, bool multi)
{
    if (sh.V == 1) return row_kernel_select_v<float, 1>(sh, fulln, flavour, multi);
    if (sh.V == 4) return row_kernel_select_v<float, 4>(sh, fulln, flavour, multi);
    return nullptr;
}
}  // namespace bsm
