"""ORACLE — TEST INFRASTRUCTURE ONLY (not the product path).

ctypes front-end for oracle/_build/liboracle.so, the plain-C restatement of the reference's
``Csr<T>::mul_dense`` (/root/reference/src/sparse.rs:426-446) and the ``Csr`` construction
rules it depends on (sparse.rs:116-132, 193-265).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module, and only as the checker or the reported CPU baseline.

Parity status: pinned on the reference's integer KATs (see tests/test_oracle_kats.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

_CT = {"i32": C.c_int32, "f32": C.c_float, "f64": C.c_double}
_NP = {"i32": np.int32, "f32": np.float32, "f64": np.float64}

MATERR = {0: None, 1: "MatrixFinalised", 2: "panic: big eek", 3: "IncorrectDimensions"}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("ref_cpu.c", "ref_cpu_impl.h", "ref_solve_band.c")
    ):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        sz, vp = C.c_size_t, C.c_void_p
        for sfx, ct in _CT.items():
            P = C.POINTER(ct)
            PP = C.POINTER(P)
            f = lambda n: getattr(L, f"{n}_{sfx}")
            f("ocsr_new").restype = vp
            f("ocsr_new").argtypes = [sz, sz, sz]
            f("ocsr_free").restype = None
            f("ocsr_free").argtypes = [vp]
            f("ocsr_insert").restype = C.c_int
            f("ocsr_insert").argtypes = [vp, ct, sz, sz]
            f("ocsr_finalise").restype = C.c_int
            f("ocsr_finalise").argtypes = [vp]
            f("ocsr_get_nnz").restype = sz
            f("ocsr_get_nnz").argtypes = [vp]
            f("ocsr_from_data").restype = vp
            f("ocsr_from_data").argtypes = [P, sz, sz]
            f("ocsr_from_raw").restype = vp
            f("ocsr_from_raw").argtypes = [sz, sz, sz, P, C.POINTER(sz), C.POINTER(sz)]
            f("ocsr_get_row_compact").restype = sz
            f("ocsr_get_row_compact").argtypes = [vp, sz, P, C.POINTER(sz), sz]
            f("ocsr_mul_dense").restype = C.c_int
            f("ocsr_mul_dense").argtypes = [vp, PP, sz, sz, C.c_int, C.POINTER(vp)]
            f("ocsr_mul_dense_rows").restype = C.c_int
            f("ocsr_mul_dense_rows").argtypes = [vp, PP, sz, sz, sz, sz, P]
            for acc in ("rows", "cols", "v_len", "row_index_len"):
                f(f"ocsr_{acc}").restype = sz
                f(f"ocsr_{acc}").argtypes = [vp]
            f("ocsr_is_finalised").restype = C.c_int
            f("ocsr_is_finalised").argtypes = [vp]
            f("ocsr_v").restype = P
            f("ocsr_v").argtypes = [vp]
            f("ocsr_col_index").restype = C.POINTER(sz)
            f("ocsr_col_index").argtypes = [vp]
            f("ocsr_row_index").restype = C.POINTER(sz)
            f("ocsr_row_index").argtypes = [vp]
        L.ocsr_mul_vector_i32.restype = C.c_int
        L.ocsr_mul_vector_i32.argtypes = [vp, C.POINTER(C.c_int32), sz, C.POINTER(C.c_int32), sz]
        for sfx, ct in (("f32", C.c_float), ("f64", C.c_double)):
            fn = getattr(L, f"ocsr_time_mul_dense_rows_{sfx}")
            fn.restype = C.c_double
            fn.argtypes = [vp, C.POINTER(C.POINTER(ct)), sz, sz, sz, sz, C.c_int, C.POINTER(sz)]
        _lib = L
    return _lib


def _sfx(dtype) -> str:
    dt = np.dtype(dtype)
    for k, v in _NP.items():
        if dt == np.dtype(v):
            return k
    raise TypeError(f"oracle supports i32/f32/f64, not {dt}")


def _colptrs(cols, sfx):
    ct = _CT[sfx]
    arr = (C.POINTER(ct) * len(cols))()
    keep = []
    for i, c in enumerate(cols):
        a = np.ascontiguousarray(c, dtype=_NP[sfx])
        keep.append(a)
        arr[i] = a.ctypes.data_as(C.POINTER(ct))
    return arr, keep


class OracleCsr:
    """Restated ``Csr<T>`` (reference: src/sparse.rs:68-78) living in the C oracle."""

    def __init__(self, handle, sfx):
        self._h = handle
        self._sfx = sfx

    # -- construction (sparse.rs:117, 193, 222, 206) ---------------------------------
    @classmethod
    def new(cls, dims, dtype=np.int32, capacity=0):
        sfx = _sfx(dtype)
        return cls(getattr(lib(), f"ocsr_new_{sfx}")(dims[0], dims[1], capacity), sfx)

    @classmethod
    def from_data(cls, rows, dtype=np.int32):
        sfx = _sfx(dtype)
        a = np.ascontiguousarray(rows, dtype=_NP[sfx])
        assert a.ndim == 2
        h = getattr(lib(), f"ocsr_from_data_{sfx}")(
            a.ctypes.data_as(C.POINTER(_CT[sfx])), a.shape[0], a.shape[1])
        return cls(h, sfx)

    @classmethod
    def from_raw(cls, dims, v, col_index, row_index):
        sfx = _sfx(np.asarray(v).dtype)
        v = np.ascontiguousarray(v, dtype=_NP[sfx])
        ci = np.ascontiguousarray(col_index, dtype=np.uint64)
        ri = np.ascontiguousarray(row_index, dtype=np.uint64)
        assert ri.shape[0] == dims[0] + 1 and ci.shape[0] == v.shape[0]
        h = getattr(lib(), f"ocsr_from_raw_{sfx}")(
            dims[0], dims[1], v.shape[0], v.ctypes.data_as(C.POINTER(_CT[sfx])),
            ci.ctypes.data_as(C.POINTER(C.c_size_t)), ri.ctypes.data_as(C.POINTER(C.c_size_t)))
        return cls(h, sfx)

    def insert(self, value, row, col):
        rc = getattr(lib(), f"ocsr_insert_{self._sfx}")(self._h, value, row, col)
        if rc:
            raise RuntimeError(MATERR[rc])

    def finalise(self):
        rc = getattr(lib(), f"ocsr_finalise_{self._sfx}")(self._h)
        if rc:
            raise RuntimeError(MATERR[rc])
        return self

    def __del__(self):
        try:
            if self._h:
                getattr(lib(), f"ocsr_free_{self._sfx}")(self._h)
                self._h = None
        except Exception:
            pass

    # -- accessors -----------------------------------------------------------------------
    def _get(self, name):
        return getattr(lib(), f"ocsr_{name}_{self._sfx}")(self._h)

    @property
    def dims(self):
        return (self._get("rows"), self._get("cols"))

    @property
    def is_finalised(self):
        return bool(self._get("is_finalised"))

    def get_nnz(self):
        return self._get("get_nnz")

    @property
    def v(self):
        n = self._get("v_len")
        return np.ctypeslib.as_array(self._get("v"), shape=(n,)).copy() if n else np.zeros(0, _NP[self._sfx])

    @property
    def col_index(self):
        n = self._get("v_len")
        return (np.ctypeslib.as_array(self._get("col_index"), shape=(n,)).astype(np.uint64)
                if n else np.zeros(0, np.uint64))

    @property
    def row_index(self):
        n = self._get("row_index_len")
        return np.ctypeslib.as_array(self._get("row_index"), shape=(n,)).astype(np.uint64)

    def raw(self):
        return self.v, self.col_index, self.row_index

    def get_row_compact(self, index):
        cap = max(1, self._get("cols"))
        vals = np.zeros(cap, _NP[self._sfx])
        cols = np.zeros(cap, np.uint64)
        n = getattr(lib(), f"ocsr_get_row_compact_{self._sfx}")(
            self._h, index, vals.ctypes.data_as(C.POINTER(_CT[self._sfx])),
            cols.ctypes.data_as(C.POINTER(C.c_size_t)), cap)
        return [(vals[i].item(), int(cols[i]), index) for i in range(n)]

    def to_dense(self):
        """Row-major densification (test helper; zeros where nothing is stored)."""
        r, c = self.dims
        out = np.zeros((r, c), _NP[self._sfx])
        v, ci, ri = self.raw()
        for i in range(r):
            for e in range(int(ri[i]), int(ri[i + 1])):
                out[i, int(ci[e])] = v[e]
        return out

    # -- the hot path ------------------------------------------------------------------
    def mul_dense(self, rhs_columns, faithful=True):
        """``Csr::mul_dense`` (sparse.rs:426-446). ``rhs_columns[c]`` is COLUMN c of the
        reference's column-major ``Dense`` (dense.rs:21-29). Returns an OracleCsr (zero-dropped,
        finalised) or raises RuntimeError('IncorrectDimensions')."""
        ncols = len(rhs_columns)
        nrows = len(rhs_columns[0]) if ncols else 0
        arr, keep = _colptrs(rhs_columns, self._sfx)
        out = C.c_void_p()
        rc = getattr(lib(), f"ocsr_mul_dense_{self._sfx}")(
            self._h, arr, nrows, ncols, int(faithful), C.byref(out))
        if rc:
            raise RuntimeError(MATERR[rc])
        return OracleCsr(out.value, self._sfx)

    def mul_dense_rows(self, rhs_columns, row_begin, row_end):
        """Same arithmetic, rows [row_begin,row_end), dense row-major output (no zero-drop)."""
        ncols = len(rhs_columns)
        nrows = len(rhs_columns[0]) if ncols else 0
        arr, keep = _colptrs(rhs_columns, self._sfx)
        out = np.zeros((row_end - row_begin, ncols), _NP[self._sfx])
        rc = getattr(lib(), f"ocsr_mul_dense_rows_{self._sfx}")(
            self._h, arr, nrows, ncols, row_begin, row_end,
            out.ctypes.data_as(C.POINTER(_CT[self._sfx])))
        if rc:
            raise RuntimeError(MATERR[rc])
        return out

    def mul_vector(self, rhs, out_len=None):
        """``Csr::mul_vector`` (sparse.rs:468-482), i32 only (KAT 1501-1529)."""
        assert self._sfx == "i32"
        rhs = np.ascontiguousarray(rhs, np.int32)
        out = np.zeros(self.dims[0] if out_len is None else out_len, np.int32)
        rc = lib().ocsr_mul_vector_i32(self._h, rhs.ctypes.data_as(C.POINTER(C.c_int32)), rhs.shape[0],
                                       out.ctypes.data_as(C.POINTER(C.c_int32)), out.shape[0])
        if rc:
            raise RuntimeError(MATERR[rc])
        return out

    def time_lean_parallel(self, rhs_columns, row_begin, row_end, threads=None):
        """NOT the reference: (seconds, threads, row-major product) of the lean multi-threaded variant (same
        arithmetic order, rows split over `threads` POSIX threads, dense output, no per-row allocation)."""
        assert self._sfx in ("f32", "f64")
        threads = int(threads or os.cpu_count() or 1)
        ncols = len(rhs_columns)
        arr, keep = _colptrs(rhs_columns, self._sfx)
        out = np.empty((row_end - row_begin, ncols), np.float64 if self._sfx == "f64" else np.float32)
        f = getattr(lib(), f"ocsr_time_lean_parallel_{self._sfx}")
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, type(arr), C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int]
        t = f(self._h, arr, ncols, row_begin, row_end, out.ctypes.data_as(C.c_void_p), threads)
        return t, threads, out

    def time_mul_dense_rows(self, rhs_columns, row_begin, row_end, faithful=True, rhs_row_count=None):
        """Seconds for the faithful multiply of rows [row_begin,row_end) (CPU baseline leg).
        ``rhs_row_count`` overrides the row count used by the dims check (sparse.rs:427-429) when the
        columns passed are only the window of B that the sampled rows reference."""
        assert self._sfx in ("f32", "f64")
        ncols = len(rhs_columns)
        nrows = len(rhs_columns[0]) if ncols else 0
        if rhs_row_count is not None:
            nrows = int(rhs_row_count)
        arr, keep = _colptrs(rhs_columns, self._sfx)
        nnz = C.c_size_t(0)
        t = getattr(lib(), f"ocsr_time_mul_dense_rows_{self._sfx}")(
            self._h, arr, nrows, ncols, row_begin, row_end, int(faithful), C.byref(nnz))
        if t < 0:
            raise RuntimeError("IncorrectDimensions")
        return t, nnz.value
