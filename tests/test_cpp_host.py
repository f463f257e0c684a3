"""The C++ host mirror of the reference API (include/bsm.hpp) over the C ABI: the reference's own
unit tests for the hot path, transcribed to C++ (tests/cpp/test_reference_kats.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
BIN = os.path.join(CPP, "_build", "test_reference_kats")


def build():
    subprocess.run(["make", "-C", CPP], check=True, capture_output=True)


def test_cpp_host_mirror_structure_kats():
    """Construction rules, zero-skip, finalise, get_row_compact, Dense layout — no GPU involved."""
    build()
    r = subprocess.run([BIN, "--host-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok: 0 failure(s)" in r.stdout


@pytest.mark.gpu
def test_cpp_reference_kats_on_gpu():
    """test_dense_mul / test_nnz / test_mul_vector of the reference through Csr<T>::mul_dense in C++."""
    build()
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok: 0 failure(s)" in r.stdout
