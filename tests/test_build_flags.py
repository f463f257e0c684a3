"""CPU: the bit-exact kernels (vector CSR and its row-block variant) must not contain fused multiply-adds.

The reference computes `value = value + (a * b)` with two roundings (src/sparse.rs:438-439); nvcc never
contracts the _rn intrinsics, but ptxas DOES contract a packed mul.rn.f32x2 feeding an add.rn.f32x2 into one
FFMA2, which is why the kernels use the packed multiply with scalar adds only. This checks the shipped SASS."""
import os
import re
import shutil
import subprocess

import pytest

from basic_sparse_matrix_b200 import _lib


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_no_fused_multiply_add_in_the_bit_exact_kernels():
    assert os.path.exists(_lib.LIB_PATH), "build the native library first (__graft_entry__.build())"
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout
    current, fused, seen = None, {}, set()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = m.group(1)
            continue
        if current and ("spmm_rows_kernel" in current or "spmm_rowblock_kernel" in current):
            seen.add(current)
            if re.search(r"\b(FFMA2?|DFMA|HFMA2)\b", line) and "HFMA2.MMA" not in line:
                fused.setdefault(current, []).append(line.strip()[:80])
    assert len(seen) > 50, "kernels not found in the SASS dump"
    # (HFMA2 / FFMA with constant operands are ptxas idioms for moving immediates; a real product has register operands)
    real = {k: [l for l in v if not re.search(r"(HFMA2|FFMA)\S* R\d+, -?RZ|, RZ, ", l)] for k, v in fused.items()}
    real = {k: v for k, v in real.items() if v}
    assert not real, {k: v[:2] for k, v in list(real.items())[:3]}
