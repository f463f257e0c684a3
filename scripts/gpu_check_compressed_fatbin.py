"""One short GPU call: load the library built with -Xfatbin -compress-all (BSM_B200_LIB points at it), run smoke() (kernels of every
translation unit but spmm_rows_f32), an f32 vector-kernel product against the oracle and the mul_dense_s KAT. No torch import."""
import json
import os
import sys
import time

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import __graft_entry__ as g
from basic_sparse_matrix_b200 import Csr, DenseS, _lib, gen, gpu
from oracle import ref_numpy

res = {"lib": _lib.LIB_PATH, "lib_bytes": os.path.getsize(_lib.LIB_PATH)}
g.smoke()
res["smoke_s"] = round(time.time() - t0, 2)
# f32 vector kernel (spmm_rows_f32.cu), bit-exact on real-valued data
A = gpu.DeviceCsr.laplacian(24, 24, 24, dtype=np.float32)
B = gpu.DeviceDense.generate(24 ** 3, 33, seed=7, mode=gen.MODE_REAL, dtype=np.float32)
C = A.mul_dense(B, algo="vector")
v, ci, ri, dims = gen.laplacian(24, 24, 24, dtype=np.float32)
want = ref_numpy.mul_dense_rowmajor(v, ci, ri, gen.dense_rows(24 ** 3, 33, 7, gen.MODE_REAL, 0.0, np.float32))
res["f32_vector_bitwise"] = bool(np.array_equal(C.to_rowmajor().view(np.uint32), want.view(np.uint32)))
# Csr::mul_dense_s on the reference's test_dense_mul operands (sparse.rs:1082-1109)
k = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")))["mul_dense"][0]
ok = True
for dt in (np.float64, np.float32):
    m = Csr.from_data(k["csr_rows"], dt)
    ok = ok and m.mul_dense_s(DenseS.from_data(k["dense_columns"], 4, 3, dt)) == Csr.from_data(k["output_rows"], dt)
res["mul_dense_s_kat"] = bool(ok)
res["kernels_launched"] = gpu.kernel_launch_count()
res["total_s"] = round(time.time() - t0, 2)
print(json.dumps(res))
with open(os.path.join(ROOT, "gpurun_out", "r2_compressed_fatbin_check.json"), "w") as f:
    json.dump(res, f, indent=1)
