"""ctypes binding of the C ABI in include/bsm.h (libbsm_b200.so, CUDA sm_100a).

There is no CPU fallback anywhere in this package: if the shared library is missing, or no
CUDA device is usable, the first call raises.  Build with ``python -c "import __graft_entry__ as
g; g.build()"`` or ``make -C basic_sparse_matrix_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .util import MatErr, MatError

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSM_B200_LIB points at another build of the same library (A/B timing of two builds on one box)
LIB_PATH = os.environ.get("BSM_B200_LIB") or os.path.join(_HERE, "lib", "libbsm_b200.so")

BSM_OK = 0
BSM_ERR_INCORRECT_DIMENSIONS = 1
BSM_ERR_NOT_FINALISED = 2
BSM_ERR_OUT_OF_BOUNDS = 3
BSM_ERR_INDEX_OVERFLOW = 4
BSM_ERR_INVALID_ARGUMENT = 5
BSM_ERR_DTYPE_MISMATCH = 6
BSM_ERR_CUDA = 7
BSM_ERR_NCCL = 8
BSM_ERR_NO_DEVICE = 9
BSM_ERR_NOT_SUPPORTED = 10

BSM_F32, BSM_F64 = 0, 1
ALGO_AUTO, ALGO_VECTOR, ALGO_MERGE, ALGO_ROWBLOCK = 0, 1, 2, 3
ALGO_NAMES = {"auto": ALGO_AUTO, "vector": ALGO_VECTOR, "merge": ALGO_MERGE, "rowblock": ALGO_ROWBLOCK}

TUNE_A_EVICT_FIRST = 0x1
TUNE_C_STREAMING = 0x2
TUNE_FUSED = 0x4
TUNE_LITERAL = 0x80000000


class BsmError(RuntimeError):
    """Any non-MatErr failure of the native library (CUDA, NCCL, bad argument, no device)."""

    def __init__(self, status, message):
        super().__init__(f"bsm status {status}: {message}")
        self.status = status


class Tuning(C.Structure):
    _fields_ = [
        ("algo", C.c_int32), ("col_tile", C.c_int32), ("rows_per_slice", C.c_int32),
        ("stages", C.c_int32), ("warps_per_cta", C.c_int32), ("ctas_per_sm", C.c_int32),
        ("merge_items", C.c_int32), ("flags", C.c_uint32), ("rows_per_warp", C.c_int32),
        ("prefer_wide_rows", C.c_int32), ("reg_flavour", C.c_int32), ("lanes_per_row", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class LaunchInfo(C.Structure):
    _fields_ = [
        ("algo", C.c_int32), ("kernels", C.c_int32), ("vec_elems", C.c_int32),
        ("lanes_per_row", C.c_int32), ("reg_tiles", C.c_int32), ("grid", C.c_int32),
        ("block", C.c_int32), ("smem_bytes", C.c_int32), ("rows_per_slice", C.c_int32),
        ("stages", C.c_int32), ("capacity", C.c_int32), ("passes", C.c_int32),
        ("merge_items", C.c_int32), ("merge_chunks", C.c_int32), ("rows_per_warp", C.c_int32),
        ("reg_flavour", C.c_int32), ("col_tile", C.c_int32), ("reserved", C.c_int32 * 1),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


def dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.float64:
        return BSM_F64
    if dt == np.float32:
        return BSM_F32
    raise TypeError(f"the GPU path computes in f32 or f64, not {dt}")


def np_dtype(code: int):
    return np.float64 if code == BSM_F64 else np.float32


def suffix(dtype) -> str:
    return "f64" if dtype_code(dtype) == BSM_F64 else "f32"


_lib = None

u64, vp, i32 = C.c_uint64, C.c_void_p, C.c_int
PV = C.POINTER(vp)


def _declare(L):
    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("bsm_abi_version", i32)
    sig("bsm_init", i32, i32)
    sig("bsm_device_count", i32, C.POINTER(i32))
    sig("bsm_set_stream", i32, vp)
    sig("bsm_sync", i32)
    sig("bsm_last_error_string", C.c_char_p)
    sig("bsm_status_string", C.c_char_p, i32)
    sig("bsm_device_info", i32, C.POINTER(i32), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
        C.POINTER(i32), C.POINTER(i32))
    for sfx in ("f32", "f64"):
        sig(f"bsm_csr_upload_{sfx}", i32, u64, u64, u64, vp, vp, vp, u64, PV)
        sig(f"bsm_csr_upload_rows_{sfx}", i32, u64, u64, vp, vp, vp, u64, u64, PV)
        sig(f"bsm_csr_download_{sfx}", i32, vp, vp, vp, vp)
        sig(f"bsm_dense_upload_{sfx}", i32, u64, u64, vp, PV)
        sig(f"bsm_dense_download_{sfx}", i32, vp, vp)
        sig(f"bsm_mul_dense_host_{sfx}", i32, u64, u64, u64, vp, vp, vp, u64, u64, u64, vp, i32,
            C.POINTER(u64), PV, PV, PV)
        sig(f"bsm_mul_dense_host_dense_{sfx}", i32, u64, u64, u64, vp, vp, vp, u64, u64, u64, vp, vp, i32)
        sig(f"bsm_mul_dense_host_into_{sfx}", i32, u64, u64, u64, vp, vp, vp, u64, u64, u64, vp, i32, u64, vp, vp, vp,
            C.POINTER(u64))
        sig(f"bsm_mul_vector_{sfx}", i32, vp, vp, u64, vp, u64)
    sig("bsm_csr_from_device", i32, i32, u64, u64, u64, vp, vp, vp, i32, PV)
    sig("bsm_csr_free", i32, vp)
    sig("bsm_csr_info", i32, vp, C.POINTER(i32), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64))
    sig("bsm_csr_stats", i32, vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64))
    sig("bsm_csr_device_ptrs", i32, vp, PV, PV, PV)
    sig("bsm_dense_alloc", i32, i32, u64, u64, PV)
    sig("bsm_dense_borrow", i32, i32, u64, u64, vp, u64, PV)
    sig("bsm_dense_free", i32, vp)
    sig("bsm_dense_zero", i32, vp)
    sig("bsm_dense_info", i32, vp, C.POINTER(i32), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), PV)
    sig("bsm_dense_download_rowmajor", i32, vp, vp)
    sig("bsm_dense_upload_rowmajor", i32, i32, u64, u64, vp, PV)
    sig("bsm_spmm", i32, vp, vp, vp, i32)
    sig("bsm_spmm_tuned", i32, vp, vp, vp, C.POINTER(Tuning))
    sig("bsm_last_launch_info", i32, C.POINTER(LaunchInfo))
    sig("bsm_line_length_of_row", C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, u64)
    sig("bsm_plan_vector", i32, i32, u64, u64, u64, C.c_uint32, u64, C.POINTER(Tuning), i32, u64, C.POINTER(LaunchInfo))
    sig("bsm_kernel_launch_count", u64)
    sig("bsm_dense_to_csr", i32, vp, PV)
    sig("bsm_dense_residual_norm", i32, vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_double))
    sig("bsm_forward_substitution", i32, vp, vp, vp)
    sig("bsm_csr_band_structure", i32, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32))
    sig("bsm_backward_substitution", i32, vp, vp, vp)
    sig("bsm_host_free", None, vp)
    sig("bsm_partition_rows", i32, vp, u64, i32, vp)
    sig("bsm_comm_unique_id", i32, C.c_char_p)
    sig("bsm_comm_init", i32, C.c_char_p, i32, i32, PV)
    sig("bsm_comm_free", i32, vp)
    sig("bsm_allgather_rows", i32, vp, vp, vp, vp)
    sig("bsm_comm_barrier", i32, vp)
    sig("bsm_spmm_scatter", i32, vp, vp, C.POINTER(vp), i32, u64, i32)
    sig("bsm_dense_ipc_export", i32, vp, C.c_char_p)
    sig("bsm_dense_ipc_open", i32, C.c_char_p, i32, u64, u64, u64, PV)
    sig("bsm_gen_dense", i32, i32, u64, u64, u64, i32, C.c_double, PV)
    sig("bsm_gen_laplacian", i32, i32, u64, u64, u64, u64, u64, PV)
    sig("bsm_gen_band", i32, i32, u64, u64, u64, u64, PV)
    sig("bsm_gen_rmat", i32, i32, i32, u64, C.c_double, C.c_double, C.c_double, u64, i32, PV)
    sig("bsm_l2_flush", i32)
    sig("bsm_phase_timers_enable", i32, i32)
    sig("bsm_phase_timers_read", i32, C.POINTER(C.c_double), i32, i32)
    sig("bsm_phase_name", C.c_char_p, i32)


def lib() -> C.CDLL:
    """The loaded native library. Raises if it has not been built — never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BsmError(-1, f"native library not built: {LIB_PATH} is missing "
                               "(run __graft_entry__.build() or make -C basic_sparse_matrix_b200/csrc); "
                               "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def check(status: int):
    """Translate a bsm_status into the reference's error vocabulary (src/util.rs:47-55)."""
    if status == BSM_OK:
        return
    msg = lib().bsm_last_error_string().decode(errors="replace")
    if status == BSM_ERR_INCORRECT_DIMENSIONS:
        raise MatError(MatErr.IncorrectDimensions, msg)
    if status == BSM_ERR_NOT_FINALISED:
        raise MatError(MatErr.MatrixNotFinalised, msg)
    if status == BSM_ERR_OUT_OF_BOUNDS:
        raise MatError(MatErr.OutOfBounds, msg)
    raise BsmError(status, msg)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(vp)


def col_ptr_array(columns):
    """void*[n] over a list of contiguous 1-D numpy arrays (the Vec<Vec<T>> of Dense)."""
    arr = (vp * len(columns))()
    for i, c in enumerate(columns):
        arr[i] = c.ctypes.data
    return arr
