"""CPU: rust/src/gpu/ffi.rs is a pure function of include/bsm.h (tools/gen_ffi.py). The image has no Rust toolchain, so the
binding cannot be compiled here; this keeps every entry point, constant and struct field of the C ABI present in it with the
header's own argument order and widths."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ffi_rs_is_what_the_header_generates():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]


def test_every_header_function_is_bound_and_every_used_binding_exists():
    header = open(os.path.join(ROOT, "include", "bsm.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    declared = set(re.findall(r"\b(bsm_\w+)\s*\(", header))
    ffi = open(os.path.join(ROOT, "rust", "src", "gpu", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (bsm_\w+)\(", ffi))
    assert declared == bound, (declared ^ bound)
    mod = open(os.path.join(ROOT, "rust", "src", "gpu", "mod.rs")).read()
    mod = re.sub(r"//[^\n]*", " ", mod)                      # identifiers in comments (e.g. `bsm_mul_dense_host_into_*`) do not count
    used = set(re.findall(r"\b(bsm_\w+)\b", mod)) - {"bsm_csr", "bsm_dense", "bsm_comm"}
    assert used <= bound, used - bound
    consts = set(re.findall(r"ffi::(BSM_\w+)", mod))
    assert consts <= set(re.findall(r"pub const (BSM_\w+)", ffi)), consts


def test_build_rs_compiles_the_same_sources_as_the_makefile():
    mk = open(os.path.join(ROOT, "basic_sparse_matrix_b200", "csrc", "Makefile")).read()
    srcs = re.search(r"^SRCS := (.*)$", mk, flags=re.M).group(1).split()
    rs = open(os.path.join(ROOT, "rust", "build.rs")).read()
    listed = re.findall(r'"(\w+\.cu)"', rs[rs.index("let sources"):rs.index("let mut objects")])
    assert sorted(srcs) == sorted(listed), (set(srcs) ^ set(listed))
    for f in srcs:
        assert os.path.exists(os.path.join(ROOT, "basic_sparse_matrix_b200", "csrc", f)), f
