"""CPU: the C-ABI shared library loads and exports every function include/bsm.h declares."""
import ctypes
import os
import re

from basic_sparse_matrix_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "bsm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)          # strip comments
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|uint64_t|uint32_t|char)\s*\*?\s*(bsm_[a-z0-9_]+)\s*\(", hdr, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    assert len(names) >= 40
    for must in ("bsm_spmm", "bsm_csr_upload_f64", "bsm_dense_upload_f64", "bsm_dense_to_csr",
                 "bsm_mul_dense_host_f64", "bsm_mul_vector_f64", "bsm_partition_rows", "bsm_allgather_rows"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the native library first (__graft_entry__.build())"
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing


def test_abi_version_and_status_strings():
    L = _lib.lib()
    assert L.bsm_abi_version() == 1
    assert L.bsm_status_string(1) == b"IncorrectDimensions"
    assert L.bsm_status_string(0) == b"ok"


def test_python_binding_covers_every_symbol():
    L = _lib.lib()
    for n in declared_functions():
        assert getattr(L, n).argtypes is not None or n in ("bsm_abi_version", "bsm_sync", "bsm_last_error_string",
                                                           "bsm_kernel_launch_count", "bsm_l2_flush"), n


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """bsm_tuning / bsm_launch_info are passed by pointer: the ctypes mirrors must have the header's
    size and field offsets (checked by compiling a probe against include/bsm.h)."""
    import subprocess
    fields_t = [k for k, _ in _lib.Tuning._fields_]
    fields_i = [k for k, _ in _lib.LaunchInfo._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "bsm.h"', 'int main(void) {',
           'printf("T %zu\\n", sizeof(bsm_tuning)); printf("I %zu\\n", sizeof(bsm_launch_info));']
    src += [f'printf("T.{k} %zu\\n", offsetof(bsm_tuning, {k}));' for k in fields_t]
    src += [f'printf("I.{k} %zu\\n", offsetof(bsm_launch_info, {k}));' for k in fields_i]
    src += ['return 0; }']
    c = tmp_path / "probe.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    assert int(got["T"]) == ctypes.sizeof(_lib.Tuning)
    assert int(got["I"]) == ctypes.sizeof(_lib.LaunchInfo)
    for k in fields_t:
        assert int(got[f"T.{k}"]) == getattr(_lib.Tuning, k).offset, k
    for k in fields_i:
        assert int(got[f"I.{k}"]) == getattr(_lib.LaunchInfo, k).offset, k
