// spmm_rows_f64.cu — f64 instantiations of the vector-CSR kernel (see spmm_rows_kernel.cuh).
#include "spmm_rows_inst.cuh"

namespace bsm {
const void *row_kernel_select_f64(Shape sh, bool fulln, int flavour, bool multi)
{
    if (sh.V == 1) return row_kernel_select_v<double, 1>(sh, fulln, flavour, multi);
    if (sh.V == 2) return row_kernel_select_v<double, 2>(sh, fulln, flavour, multi);
    return nullptr;
}
}  // namespace bsm
