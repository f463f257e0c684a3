"""CPU: what the shipped SASS of libbsm_b200.so must contain (and must not).

1. The bit-exact kernels (vector CSR and its row-block variant) must not contain fused multiply-adds.
   The reference computes `value = value + (a * b)` with two roundings (src/sparse.rs:438-439); nvcc never
   contracts the _rn intrinsics, but ptxas DOES contract a packed mul.rn.f32x2 feeding an add.rn.f32x2 into one
   FFMA2. The row-block kernel therefore forms its f32 products as fma.rn.f32x2(a, b, -0.0) (== rn(a*b)) followed
   by add.rn.f32x2: in its SASS every FFMA2 must be paired with an FADD2 and there must be no scalar FFMA; the
   vector kernel must contain no fused multiply-add at all. The opt-in BSM_TUNE_FUSED variants of the row-block
   kernel (last template argument true) are the exception: they must contain FMAs and nothing unfused.
2. Every SpMM kernel stages its A stream with TMA bulk copies (UBLKCP) and waits on mbarriers (SYNCS); the 16-byte
   lane shapes gather B with 128-bit loads (LDG.E.128); the library is built for sm_100a only.
3. The size of the instantiation table stays bounded."""
import os
import re
import shutil
import subprocess

import pytest

from basic_sparse_matrix_b200 import _lib

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def sass():
    assert os.path.exists(_lib.LIB_PATH), "build the native library first (__graft_entry__.build())"
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout
    kernels, current = {}, None
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = m.group(1)
            kernels[current] = {}
            continue
        if current is None:
            continue
        op = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if op:
            name = op.group(1)
            for key in (name, name.split(".")[0]):
                kernels[current][key] = kernels[current].get(key, 0) + 1
    return kernels, arch


def is_fused_rowblock(name):
    # spmm_rowblock_kernel<T, V, G, RB, FULLN, FUSED, NT>: the mangled name ends ...Lb<FULLN>ELb<FUSED>ELi<NT>EEEv...
    return "spmm_rowblock_kernel" in name and re.search(r"Lb[01]ELb1ELi\dEEEv", name) is not None


def test_no_fused_multiply_add_in_the_bit_exact_kernels(sass):
    kernels, _ = sass
    exact = {k: v for k, v in kernels.items() if ("spmm_rows_kernel" in k or "spmm_rowblock_kernel" in k) and not is_fused_rowblock(k)}
    assert len(exact) > 50, "kernels not found in the SASS dump"
    packed = 0
    for k, ops in exact.items():
        assert not ops.get("FFMA") and not ops.get("DFMA"), (k, "fused multiply-add in a bit-exact kernel")
        if "spmm_rowblock_kernel" in k:
            assert ops.get("FFMA2", 0) == ops.get("FADD2", 0), (k, ops.get("FFMA2"), ops.get("FADD2"))   # product-forming FFMA2 + its FADD2
            packed += ops.get("FFMA2", 0)
        else:
            assert not ops.get("FFMA2"), (k, "packed FMA in the vector kernel")
    assert packed, "the f32 row-block kernels should use the packed f32x2 pipe"


def test_opt_in_fused_rowblock_variants_are_fused(sass):
    kernels, _ = sass
    fused = {k: v for k, v in kernels.items() if is_fused_rowblock(k)}
    assert fused, "BSM_TUNE_FUSED variants of the row-block kernel are missing"
    for k, ops in fused.items():
        assert ops.get("DFMA") or ops.get("FFMA2") or ops.get("FFMA"), (k, "no FMA in a fused variant")
        assert not ops.get("FADD2") and not ops.get("DADD") and not ops.get("DMUL"), (k, "unfused arithmetic in a fused variant")


def test_spmm_kernels_use_tma_bulk_copies_and_wide_loads(sass):
    kernels, arch = sass
    assert arch == {"sm_100a"}, arch
    spmm = {k: v for k, v in kernels.items() if re.search(r"spmm_(rows|merge|rowblock)_kernel", k)}
    assert spmm
    for k, ops in spmm.items():
        assert ops.get("UBLKCP"), (k, "no TMA bulk copy (cp.async.bulk -> UBLKCP)")
        assert ops.get("SYNCS"), (k, "no mbarrier operations")
    # 16-byte lanes: spmm_*_kernel<double, 2, ...> / <float, 4, ...> gather B rows with LDG.E.128
    wide = {k: v for k, v in spmm.items() if re.search(r"kernelI(dLi2E|fLi4E)", k)}
    assert len(wide) > 40
    for k, ops in wide.items():
        assert any(name.startswith("LDG.E.128") for name in ops), (k, "no 128-bit global loads")


def test_instantiation_table_stays_bounded(sass):
    kernels, _ = sass
    assert len(kernels) < 360, len(kernels)   # 337 today: 308 SpMM / conversion / generator kernels + 12 band substitution + 12 two-tile row-block + ...
    assert os.path.getsize(_lib.LIB_PATH) < 10 * 1024 * 1024   # 9.5 MB today: line info included, SASS stored compressed (-Xfatbin -compress-all)
