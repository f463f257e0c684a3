"""CPU: the bit-exact kernels (vector CSR and its row-block variant) must not contain fused multiply-adds.

The reference computes `value = value + (a * b)` with two roundings (src/sparse.rs:438-439); nvcc never
contracts the _rn intrinsics, but ptxas DOES contract a packed mul.rn.f32x2 feeding an add.rn.f32x2 into one
FFMA2. The row-block kernel therefore forms its f32 products as fma.rn.f32x2(a, b, -0.0) (== rn(a*b)) followed
by add.rn.f32x2: in its SASS every FFMA2 must be paired with an FADD2 and there must be no scalar FFMA; the
vector kernel must contain no fused multiply-add at all. This checks the shipped SASS."""
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
    current, fused, seen, packed = None, {}, set(), {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = m.group(1)
            continue
        if not current or not ("spmm_rows_kernel" in current or "spmm_rowblock_kernel" in current):
            continue
        seen.add(current)
        op = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not op:
            continue
        name = op.group(1).split(".")[0]
        if name in ("FFMA2", "FADD2") and "spmm_rowblock_kernel" in current:
            packed.setdefault(current, {"FFMA2": 0, "FADD2": 0})[name] += 1
        elif name in ("FFMA", "FFMA2", "DFMA"):
            fused.setdefault(current, []).append(line.strip()[:80])
    assert len(seen) > 50, "kernels not found in the SASS dump"
    assert not fused, {k: v[:2] for k, v in list(fused.items())[:3]}
    assert packed, "the f32 row-block kernels should use the packed f32x2 pipe"
    for k, c in packed.items():
        assert c["FFMA2"] == c["FADD2"], (k, c)   # every product-forming FFMA2 is followed by its separate FADD2
