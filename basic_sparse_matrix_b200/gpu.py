"""``gpu`` module: device-resident ``DeviceCsr`` / ``DeviceDense`` handles over the C ABI.

This is the Python spelling of the Rust ``gpu`` module described in INTEGRATION.md: RAII handles
that own HBM buffers, ``mul_dense`` that keeps operands and product on the device between calls,
``into_csr`` for the reference's zero-dropped result type, and the row-partitioned multi-GPU
driver (one process per GPU, B replicated, optional NCCL all-gather of C row blocks).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import LaunchInfo, Tuning, check, lib
from .dense import Dense
from .sparse import Csr
from .util import MatDim, MatErr, MatError


def init(device: int = 0) -> None:
    check(lib().bsm_init(device))


def device_count() -> int:
    n = C.c_int(0)
    st = lib().bsm_device_count(C.byref(n))
    return n.value if st == 0 else 0


def set_stream(cuda_stream_ptr) -> None:
    """Adopt an external CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
    check(lib().bsm_set_stream(C.c_void_p(cuda_stream_ptr or 0)))


def sync() -> None:
    check(lib().bsm_sync())


def l2_flush() -> None:
    check(lib().bsm_l2_flush())


def fill_zero(d: "DeviceDense") -> None:
    """Zero a device-resident dense matrix on the library stream."""
    check(lib().bsm_dense_zero(d.handle))


def kernel_launch_count() -> int:
    return int(lib().bsm_kernel_launch_count())


def last_launch_info() -> dict:
    info = LaunchInfo()
    check(lib().bsm_last_launch_info(C.byref(info)))
    return info.as_dict()


def phase_timers(enable: bool | None = None, reset: bool = True) -> dict:
    """Where the time of the host-to-host calls went since the last reset (seconds per phase); ``enable`` switches
    the timers on or off first."""
    if enable is not None:
        check(lib().bsm_phase_timers_enable(1 if enable else 0))
    buf = (C.c_double * 16)()
    check(lib().bsm_phase_timers_read(buf, 16, 1 if reset else 0))
    out = {}
    for i in range(16):
        name = lib().bsm_phase_name(i).decode()
        if name:
            out[name] = buf[i]
    return out


def device_info() -> dict:
    sm, cc1, cc2 = C.c_int(0), C.c_int(0), C.c_int(0)
    l2, hbm = C.c_size_t(0), C.c_size_t(0)
    check(lib().bsm_device_info(C.byref(sm), C.byref(l2), C.byref(hbm), C.byref(cc1), C.byref(cc2)))
    return {"sm_count": sm.value, "l2_bytes": l2.value, "hbm_bytes": hbm.value, "cc": (cc1.value, cc2.value)}


def make_tuning(algo="auto", **kw) -> Tuning:
    t = Tuning()
    t.algo = _lib.ALGO_NAMES[algo] if isinstance(algo, str) else int(algo)
    for k, v in kw.items():
        if not hasattr(t, k):
            raise TypeError(f"unknown tuning field {k}")
        setattr(t, k, v)
    return t


class _Handle:
    _free = None

    def __init__(self, handle):
        self.handle = C.c_void_p(handle)

    def close(self):
        if self.handle:
            try:
                getattr(lib(), self._free)(self.handle)
            finally:
                self.handle = C.c_void_p(None)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceDense(_Handle):
    """Row-major dense matrix in HBM (the device twin of the reference's column-major ``Dense``)."""
    _free = "bsm_dense_free"

    @classmethod
    def from_host(cls, d: Dense) -> "DeviceDense":
        sfx = _lib.suffix(d.dtype)
        cols = [np.ascontiguousarray(c) for c in d.data]
        out = C.c_void_p()
        check(getattr(lib(), f"bsm_dense_upload_{sfx}")(d.row_count, d.col_count, _lib.col_ptr_array(cols), C.byref(out)))
        return cls(out.value)

    @classmethod
    def from_rowmajor(cls, a: np.ndarray) -> "DeviceDense":
        a = np.ascontiguousarray(a)
        if a.ndim == 1:
            a = a[:, None]
        out = C.c_void_p()
        check(lib().bsm_dense_upload_rowmajor(_lib.dtype_code(a.dtype), a.shape[0], a.shape[1], _lib.ptr(a), C.byref(out)))
        return cls(out.value)

    @classmethod
    def alloc(cls, rows: int, cols: int, dtype=np.float64) -> "DeviceDense":
        out = C.c_void_p()
        check(lib().bsm_dense_alloc(_lib.dtype_code(dtype), rows, cols, C.byref(out)))
        return cls(out.value)

    @classmethod
    def borrow(cls, device_ptr: int, rows: int, cols: int, ld: int, dtype) -> "DeviceDense":
        out = C.c_void_p()
        check(lib().bsm_dense_borrow(_lib.dtype_code(dtype), rows, cols, C.c_void_p(device_ptr), ld, C.byref(out)))
        return cls(out.value)

    @classmethod
    def generate(cls, rows: int, cols: int, seed: int, mode: int, offset: float = 0.0, dtype=np.float64) -> "DeviceDense":
        out = C.c_void_p()
        check(lib().bsm_gen_dense(_lib.dtype_code(dtype), rows, cols, seed, mode, offset, C.byref(out)))
        return cls(out.value)

    def info(self) -> dict:
        dt, r, c, ld, p = C.c_int(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_void_p()
        check(lib().bsm_dense_info(self.handle, C.byref(dt), C.byref(r), C.byref(c), C.byref(ld), C.byref(p)))
        return {"dtype": _lib.np_dtype(dt.value), "rows": r.value, "cols": c.value, "ld": ld.value, "ptr": p.value or 0}

    def get_dims(self) -> MatDim:
        i = self.info()
        return MatDim(i["rows"], i["cols"])

    def to_rowmajor(self) -> np.ndarray:
        i = self.info()
        out = np.empty((i["rows"], i["cols"]), dtype=i["dtype"])
        check(lib().bsm_dense_download_rowmajor(self.handle, _lib.ptr(out)))
        return out

    def to_host(self, into: Dense | None = None) -> Dense:
        """Download into the reference's column-major ``Dense`` (device-side transpose)."""
        i = self.info()
        if into is None:
            into = Dense.new_default_with_dims(i["cols"], i["rows"], i["dtype"])
        sfx = _lib.suffix(i["dtype"])
        check(getattr(lib(), f"bsm_dense_download_{sfx}")(self.handle, _lib.col_ptr_array(into.data)))
        return into

    def residual_norm(self, b: "DeviceDense") -> tuple:
        """(||self - b||_F, ||b||_F) — the residual check of BASELINE config 5 with self = A·X."""
        r, n = C.c_double(0.0), C.c_double(0.0)
        check(lib().bsm_dense_residual_norm(self.handle, b.handle, C.byref(r), C.byref(n)))
        return r.value, n.value

    def ipc_export(self) -> bytes:
        """64 opaque bytes another process of the same box can open with ``DeviceDense.ipc_open``."""
        buf = C.create_string_buffer(64)
        check(lib().bsm_dense_ipc_export(self.handle, buf))
        return buf.raw

    @classmethod
    def ipc_open(cls, handle: bytes, rows: int, cols: int, ld: int, dtype) -> "DeviceDense":
        out = C.c_void_p()
        check(lib().bsm_dense_ipc_open(handle, _lib.dtype_code(dtype), rows, cols, ld, C.byref(out)))
        return cls(out.value)

    def into_csr(self) -> "DeviceCsr":
        """Zero-dropping compaction = the reference's result construction (sparse.rs:442, 222-233,
        206-219), on the device."""
        out = C.c_void_p()
        check(lib().bsm_dense_to_csr(self.handle, C.byref(out)))
        return DeviceCsr(out.value)


class DeviceCsr(_Handle):
    """CSR operand in HBM: values, u32 column indices, u32 row pointer."""
    _free = "bsm_csr_free"

    @classmethod
    def from_host(cls, m: Csr, row_begin: int | None = None, row_end: int | None = None) -> "DeviceCsr":
        sfx = _lib.suffix(m.dtype)
        v, ci, ri = m.raw_parts()
        out = C.c_void_p()
        if row_begin is None:
            check(getattr(lib(), f"bsm_csr_upload_{sfx}")(
                m.dims.rows, m.dims.cols, len(v), _lib.ptr(v), _lib.ptr(ci), _lib.ptr(ri), len(ri), C.byref(out)))
        else:
            if not m.is_finalised:
                raise MatError(MatErr.MatrixNotFinalised)
            check(getattr(lib(), f"bsm_csr_upload_rows_{sfx}")(
                m.dims.rows, m.dims.cols, _lib.ptr(v), _lib.ptr(ci), _lib.ptr(ri), row_begin, row_end, C.byref(out)))
        return cls(out.value)

    @classmethod
    def laplacian(cls, nx, ny, nz=1, row_begin=0, row_end=None, dtype=np.float64) -> "DeviceCsr":
        out = C.c_void_p()
        n = nx * ny * nz
        check(lib().bsm_gen_laplacian(_lib.dtype_code(dtype), nx, ny, nz, row_begin, n if row_end is None else row_end, C.byref(out)))
        return cls(out.value)

    @classmethod
    def band(cls, n, hb, row_begin=0, row_end=None, dtype=np.float64) -> "DeviceCsr":
        out = C.c_void_p()
        check(lib().bsm_gen_band(_lib.dtype_code(dtype), n, hb, row_begin, n if row_end is None else row_end, C.byref(out)))
        return cls(out.value)

    @classmethod
    def rmat(cls, scale, edges, a=0.57, b=0.19, c=0.19, seed=3, mode=0, dtype=np.float64) -> "DeviceCsr":
        out = C.c_void_p()
        check(lib().bsm_gen_rmat(_lib.dtype_code(dtype), scale, edges, a, b, c, seed, mode, C.byref(out)))
        return cls(out.value)

    def info(self) -> dict:
        dt = C.c_int(0)
        r, c, nnz, mx = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        check(lib().bsm_csr_info(self.handle, C.byref(dt), C.byref(r), C.byref(c), C.byref(nnz), C.byref(mx)))
        return {"dtype": _lib.np_dtype(dt.value), "rows": r.value, "cols": c.value, "nnz": nnz.value, "max_row_nnz": mx.value}

    def band_structure(self) -> dict:
        """Half-bandwidth when the matrix is a proper lower / upper band factor (what the substitutions specialise on), else -1."""
        lo, up = C.c_int32(-1), C.c_int32(-1)
        check(lib().bsm_csr_band_structure(self.handle, C.byref(lo), C.byref(up)))
        return {"lower_hb": lo.value, "upper_hb": up.value}

    def stats(self) -> dict:
        """What the upload measured over all rows: longest row, column range, majority stencil line length (0 = none)."""
        mx, lo, hi, ll = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        check(lib().bsm_csr_stats(self.handle, C.byref(mx), C.byref(lo), C.byref(hi), C.byref(ll)))
        return {"max_row_nnz": mx.value, "col_min": lo.value, "col_max": hi.value, "line_length": ll.value}

    def get_dims(self) -> MatDim:
        i = self.info()
        return MatDim(i["rows"], i["cols"])

    def get_nnz(self) -> int:
        return self.info()["nnz"]

    def to_host(self) -> Csr:
        i = self.info()
        v = np.empty(i["nnz"], dtype=i["dtype"])
        ci = np.empty(i["nnz"], dtype=np.uint64)
        ri = np.empty(i["rows"] + 1, dtype=np.uint64)
        sfx = _lib.suffix(i["dtype"])
        check(getattr(lib(), f"bsm_csr_download_{sfx}")(self.handle, _lib.ptr(v), _lib.ptr(ci), _lib.ptr(ri)))
        return Csr.from_raw_parts((i["rows"], i["cols"]), v, ci, ri)

    def mul_dense(self, rhs: DeviceDense, out: DeviceDense | None = None, algo="auto", tuning: Tuning | None = None) -> DeviceDense:
        """Device-resident ``mul_dense`` (sparse.rs:426-446): returns the dense product in HBM;
        ``.into_csr()`` gives the reference's zero-dropped result type."""
        if out is None:
            ai, bi = self.info(), rhs.info()
            if ai["cols"] != bi["rows"]:
                raise MatError(MatErr.IncorrectDimensions)
            out = DeviceDense.alloc(ai["rows"], bi["cols"], ai["dtype"])
        if tuning is not None:
            check(lib().bsm_spmm_tuned(self.handle, rhs.handle, out.handle, C.byref(tuning)))
        else:
            check(lib().bsm_spmm(self.handle, rhs.handle, out.handle, _lib.ALGO_NAMES[algo] if isinstance(algo, str) else algo))
        return out


    def forward_substitution(self, b: DeviceDense, out: DeviceDense | None = None) -> DeviceDense:
        """``forward_substitution(l, b)`` of the reference's ``solve`` (lib.rs:28-46) with ``self`` = L: solves L y = b for
        every column of ``b`` on the device, bit-identical to the reference's sequential loops."""
        if out is None:
            bi = b.info()
            out = DeviceDense.alloc(bi["rows"], bi["cols"], bi["dtype"])
        check(lib().bsm_forward_substitution(self.handle, b.handle, out.handle))
        return out

    def backward_substitution(self, y: DeviceDense, out: DeviceDense | None = None) -> DeviceDense:
        """``backward_substitution(l_star, y)`` (lib.rs:49-65) with ``self`` = L* (the transposed factor, diagonal first in
        every row): solves L* x = y on the device."""
        if out is None:
            yi = y.info()
            out = DeviceDense.alloc(yi["rows"], yi["cols"], yi["dtype"])
        check(lib().bsm_backward_substitution(self.handle, y.handle, out.handle))
        return out

    def mul_dense_scatter(self, rhs: DeviceDense, full_buffers, row_offset: int, algo="auto") -> None:
        """Fused multiply + all-gather: every finished C row goes to ``full_buffers[0]`` (this rank's full
        result) and to all the others (peer GPUs' full results, mapped with ``ipc_open``), at global row
        ``row_offset + local row``. No collective; barrier across ranks before reading."""
        arr = (C.c_void_p * len(full_buffers))(*[f.handle for f in full_buffers])
        check(lib().bsm_spmm_scatter(self.handle, rhs.handle, arr, len(full_buffers), row_offset,
                                     _lib.ALGO_NAMES[algo] if isinstance(algo, str) else algo))


# ---- row-partitioned multi-GPU ------------------------------------------------------------------
def partition_rows(row_index: np.ndarray, parts: int) -> np.ndarray:
    """nnz-balanced contiguous row split (host): ``bounds[p]`` = first row of part ``p``."""
    ri = np.ascontiguousarray(row_index, dtype=np.uint64)
    bounds = np.zeros(parts + 1, dtype=np.uint64)
    check(lib().bsm_partition_rows(_lib.ptr(ri), len(ri) - 1, parts, _lib.ptr(bounds)))
    return bounds


class Comm(_Handle):
    """NCCL communicator for the optional all-gather of C row blocks."""
    _free = "bsm_comm_free"

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().bsm_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def init(cls, unique_id: bytes, nranks: int, rank: int) -> "Comm":
        out = C.c_void_p()
        check(lib().bsm_comm_init(unique_id, nranks, rank, C.byref(out)))
        return cls(out.value)

    def barrier(self) -> None:
        check(lib().bsm_comm_barrier(self.handle))

    def allgather_rows(self, local_block: DeviceDense, bounds: np.ndarray, full: DeviceDense) -> None:
        b = np.ascontiguousarray(bounds, dtype=np.uint64)
        check(lib().bsm_allgather_rows(self.handle, local_block.handle, _lib.ptr(b), full.handle))
