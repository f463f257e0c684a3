"""Host ``Csr<T>`` with the reference's construction rules and the GPU-backed ``mul_dense``.

Mirrors /root/reference/src/sparse.rs for the hot-path surface only: ``new`` /
``new_with_capacity`` (116-132), ``from_data`` (193-203), ``insert`` (222-233) ->
``insert_unchecked`` (237-250), ``finalise`` (206-219), ``get_nnz`` / ``get_density`` (162-168),
``get_row_compact`` (252-265), ``GetDims`` (418-422), and the three "x dense" entry points
``mul_dense`` (426-446), ``mul_vector`` (468-482).  The multiplications run on the B200 through the
C ABI (include/bsm.h); there is no CPU implementation of them in this package.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .dense import Dense
from .util import GetDims, MatDim, MatErr, MatError


@dataclass
class CsrEntry:
    """``CsrEntry { v, col_index, row_index }`` (sparse.rs:80-85)."""
    v: object
    col_index: int
    row_index: int


class Csr(GetDims):
    """CSR matrix: ``v`` (values), ``col_index`` (usize), ``row_index`` (usize row pointer of
    length rows+1 once finalised) — sparse.rs:68-78."""

    def __init__(self, dims, dtype=np.float64, capacity: int = 0):
        self.dims = MatDim.of(dims)
        self.dtype = np.dtype(dtype)
        self._v = []
        self._col_index = []
        self._row_index = [0]                 # row_index: vec![0]   sparse.rs:127
        self.is_finalised = False
        self.iter_v_index = 0
        self.iter_row_index = 0
        self._frozen = None                   # numpy copies of the three arrays once finalised

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def new(cls, dims, dtype=np.float64) -> "Csr":                       # sparse.rs:117-119
        return cls(dims, dtype, 0)

    @classmethod
    def new_with_capacity(cls, dims, capacity: int, dtype=np.float64) -> "Csr":   # sparse.rs:121-132
        return cls(dims, dtype, capacity)

    @classmethod
    def from_data(cls, data, dtype=np.float64) -> "Csr":                 # sparse.rs:193-203 (data[r] is ROW r)
        rows = len(data)
        cols = len(data[0])
        m = cls((rows, cols), dtype)
        for i, row in enumerate(data):
            for j, val in enumerate(row):
                m.insert(val, i, j)
        return m.finalise()

    @classmethod
    def from_raw_parts(cls, dims, v, col_index, row_index, is_finalised=True) -> "Csr":
        """Adopt existing arrays (the reference keeps its fields private, sparse.rs:69-78; a
        drop-in crate adds the equivalent ``pub(crate)`` constructor for the ``gpu`` module)."""
        v = np.ascontiguousarray(v)
        m = cls(dims, v.dtype)
        m._frozen = (v, np.ascontiguousarray(col_index, dtype=np.uint64),
                     np.ascontiguousarray(row_index, dtype=np.uint64))
        m._v = m._col_index = m._row_index = None
        m.is_finalised = bool(is_finalised)
        return m

    def insert(self, value, row: int, col: int) -> None:                 # sparse.rs:222-233
        if self.is_finalised:
            raise MatError(MatErr.MatrixFinalised)
        value = self.dtype.type(value)
        if value != self.dtype.type(0):        # T::default() is skipped (-0.0 too; NaN is kept)
            self._insert_unchecked(value, row, col)

    def _insert_unchecked(self, value, row: int, col: int) -> None:      # sparse.rs:237-250
        self._v.append(value)
        self._col_index.append(col)
        ri = self._row_index
        if row > len(ri) - 1:
            if row > len(ri):
                ri.append(len(self._v) - 1)
                for _ in range(len(ri), row + 1):
                    ri.append(ri[-1])
            else:
                ri.append(len(self._v) - 1)

    def finalise(self) -> "Csr":                                        # sparse.rs:206-219
        if not self.is_finalised:
            self.is_finalised = True
            if self.dims.rows < len(self._row_index):
                raise RuntimeError("big eek")                            # panic!("big eek")
            required_spacers = self.dims.rows - len(self._row_index)
            nnz = len(self._v)
            self._row_index.extend([nnz] * required_spacers)
            self._row_index.append(nnz)
            self._frozen = (np.array(self._v, dtype=self.dtype),
                            np.array(self._col_index, dtype=np.uint64),
                            np.array(self._row_index, dtype=np.uint64))
            self._v = self._col_index = self._row_index = None
        return self

    # ---- raw views ----------------------------------------------------------------------
    @property
    def v(self) -> np.ndarray:
        return self._frozen[0] if self._frozen is not None else np.array(self._v, dtype=self.dtype)

    @property
    def col_index(self) -> np.ndarray:
        return self._frozen[1] if self._frozen is not None else np.array(self._col_index, dtype=np.uint64)

    @property
    def row_index(self) -> np.ndarray:
        return self._frozen[2] if self._frozen is not None else np.array(self._row_index, dtype=np.uint64)

    def raw_parts(self):
        return self.v, self.col_index, self.row_index

    def get_dims(self) -> MatDim:                                        # sparse.rs:418-422
        return self.dims

    def get_nnz(self) -> int:                                            # sparse.rs:162-164
        ri = self.row_index
        return int(ri[-1]) if len(ri) else 0

    def get_density(self) -> float:                                      # sparse.rs:166-168
        return float(np.float32(len(self.v)) / np.float32(self.dims.rows * self.dims.cols))

    def get_row_compact(self, index: int):                               # sparse.rs:252-265
        v, ci, ri = self.raw_parts()
        row_start = int(ri[index])
        row_end = len(v) if index == len(ri) - 1 else int(ri[index + 1])
        return [CsrEntry(v[e], int(ci[e]), index) for e in range(row_start, row_end)]

    def to_dense_rowmajor(self) -> np.ndarray:
        """Densified copy (test helper): zeros where nothing is stored; later duplicates win."""
        v, ci, ri = self.raw_parts()
        out = np.zeros((self.dims.rows, self.dims.cols), dtype=self.dtype)
        rows = np.repeat(np.arange(self.dims.rows), np.diff(ri.astype(np.int64)))
        out[rows, ci.astype(np.int64)] = v
        return out

    def __eq__(self, other):                                             # #[derive(PartialEq)] sparse.rs:68
        if not isinstance(other, Csr):
            return NotImplemented
        return (self.dims == other.dims and self.is_finalised == other.is_finalised
                and self.iter_v_index == other.iter_v_index and self.iter_row_index == other.iter_row_index
                and np.array_equal(self.v, other.v) and np.array_equal(self.col_index, other.col_index)
                and np.array_equal(self.row_index, other.row_index))

    def __repr__(self):
        return f"Csr(dims={self.dims}, nnz={len(self.v)}, finalised={self.is_finalised})"

    # ---- the hot path (GPU) ------------------------------------------------------------------
    def _check_multipliable(self):
        if not self.is_finalised:
            # the reference indexes row_index[row+1] and panics when it is shorter than rows+1
            raise MatError(MatErr.MatrixNotFinalised, "finalise() the matrix before multiplying")
        _lib.dtype_code(self.dtype)

    def mul_dense(self, rhs: Dense, algo: str = "auto") -> "Csr":
        """``Csr::mul_dense(&self, rhs:&Dense<T>) -> Result<Csr<T>,MatErr>`` (sparse.rs:426-446).

        Runs on the GPU: A and B are uploaded, multiplied by the sm_100a kernels, every output is
        passed through the zero-dropping ``insert`` (device-side compaction) and the finalised
        result ``Csr`` comes back in the reference layout.  Raises
        ``MatError(MatErr.IncorrectDimensions)`` when ``self.cols != rhs.rows`` (sparse.rs:427-429)."""
        if self.dims.cols != rhs.get_dims().rows:
            raise MatError(MatErr.IncorrectDimensions)
        self._check_multipliable()
        if rhs.dtype != self.dtype:
            raise TypeError("Csr and Dense must have the same element type")
        sfx = _lib.suffix(self.dtype)
        L = _lib.lib()
        v, ci, ri = self.raw_parts()
        cols = [np.ascontiguousarray(c) for c in rhs.data]
        cp = _lib.col_ptr_array(cols)
        out_nnz = C.c_uint64(0)
        ov, oc, orow = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(getattr(L, f"bsm_mul_dense_host_{sfx}")(
            self.dims.rows, self.dims.cols, len(v), _lib.ptr(v), _lib.ptr(ci), _lib.ptr(ri), len(ri),
            rhs.row_count, rhs.col_count, cp, _lib.ALGO_NAMES[algo], C.byref(out_nnz),
            C.byref(ov), C.byref(oc), C.byref(orow)))
        try:
            nnz = out_nnz.value
            ct = C.c_double if sfx == "f64" else C.c_float
            rv = np.ctypeslib.as_array(C.cast(ov, C.POINTER(ct)), shape=(max(nnz, 1),))[:nnz].copy()
            rc = np.ctypeslib.as_array(C.cast(oc, C.POINTER(C.c_uint64)), shape=(max(nnz, 1),))[:nnz].copy()
            rr = np.ctypeslib.as_array(C.cast(orow, C.POINTER(C.c_uint64)), shape=(self.dims.rows + 1,)).copy()
        finally:
            L.bsm_host_free(ov)
            L.bsm_host_free(oc)
            L.bsm_host_free(orow)
        return Csr.from_raw_parts((self.dims.rows, rhs.col_count), rv, rc, rr)

    def mul_dense_s(self, rhs, algo: str = "auto") -> "Csr":
        """``Csr::mul_dense_s(&self, rhs:&DenseS<T,ROWS,COLS>) -> Result<Csr<T>,MatErr>`` (sparse.rs:448-466): the same
        loop nest as ``mul_dense`` against the fixed-size operand, so the same GPU call.  The dimension check is against
        the const parameter ``ROWS`` (sparse.rs:449) and the result has ``rhs.get_dims().cols`` columns (sparse.rs:450,453);
        a ``DenseS`` whose recorded ``col_count`` exceeds ``COLS`` (dense_static.rs:21-35) is the reference's index panic,
        here an ``IndexError`` from ``get_col``."""
        if self.dims.cols != rhs.ROWS:
            raise MatError(MatErr.IncorrectDimensions)
        n = rhs.get_dims().cols
        return self.mul_dense(Dense(n, rhs.ROWS, [rhs.get_col(c) for c in range(n)]), algo=algo)

    def mul_dense_csr_into(self, rhs: Dense, out_v: np.ndarray, out_col_index: np.ndarray, out_row_index: np.ndarray,
                           algo: str = "auto") -> "Csr":
        """The literal ``mul_dense`` into caller-provided result arrays (``bsm_mul_dense_host_into_*``): ``out_v`` /
        ``out_col_index`` hold up to ``len(out_v)`` entries, ``out_row_index`` rows+1.  Returns a ``Csr`` VIEW of the arrays
        (no copy).  With pinned arrays the device->host copies overlap the computation; this is the end-to-end call
        ``bench.py`` times."""
        if self.dims.cols != rhs.get_dims().rows:
            raise MatError(MatErr.IncorrectDimensions)
        self._check_multipliable()
        if rhs.dtype != self.dtype or out_v.dtype != self.dtype:
            raise TypeError("Csr, Dense and the result values must have the same element type")
        if out_col_index.dtype != np.uint64 or out_row_index.dtype != np.uint64 or len(out_row_index) < self.dims.rows + 1:
            raise TypeError("result indices are usize (uint64); row_index needs rows+1 entries")
        sfx = _lib.suffix(self.dtype)
        v, ci, ri = self.raw_parts()
        cols = [np.ascontiguousarray(c) for c in rhs.data]
        out_nnz = C.c_uint64(0)
        _lib.check(getattr(_lib.lib(), f"bsm_mul_dense_host_into_{sfx}")(
            self.dims.rows, self.dims.cols, len(v), _lib.ptr(v), _lib.ptr(ci), _lib.ptr(ri), len(ri),
            rhs.row_count, rhs.col_count, _lib.col_ptr_array(cols), _lib.ALGO_NAMES[algo],
            min(len(out_v), len(out_col_index)), _lib.ptr(out_v), _lib.ptr(out_col_index), _lib.ptr(out_row_index), C.byref(out_nnz)))
        nnz = out_nnz.value
        return Csr.from_raw_parts((self.dims.rows, rhs.col_count), out_v[:nnz], out_col_index[:nnz], out_row_index[:self.dims.rows + 1])

    def mul_dense_into(self, rhs: Dense, out: Dense | None = None, algo: str = "auto") -> Dense:
        """Same product as ``mul_dense`` with a DENSE result in the reference's column-major layout
        (no zero-drop): host ``Csr`` and host ``Dense`` in, host ``Dense`` out, through the pipelined
        C-ABI call ``bsm_mul_dense_host_dense_*`` (chunks of B rows in, blocks of output rows out, overlapped).
        ``bench.py`` reports it as ``e2e_dense``; pass pinned column buffers for full overlap."""
        if self.dims.cols != rhs.get_dims().rows:
            raise MatError(MatErr.IncorrectDimensions)
        self._check_multipliable()
        if rhs.dtype != self.dtype:
            raise TypeError("Csr and Dense must have the same element type")
        if out is None:
            out = Dense.new_default_with_dims(rhs.col_count, self.dims.rows, self.dtype)
        sfx = _lib.suffix(self.dtype)
        v, ci, ri = self.raw_parts()
        cols = [np.ascontiguousarray(c) for c in rhs.data]
        _lib.check(getattr(_lib.lib(), f"bsm_mul_dense_host_dense_{sfx}")(
            self.dims.rows, self.dims.cols, len(v), _lib.ptr(v), _lib.ptr(ci), _lib.ptr(ri), len(ri),
            rhs.row_count, rhs.col_count, _lib.col_ptr_array(cols), _lib.col_ptr_array(out.data), _lib.ALGO_NAMES[algo]))
        return out

    def mul_vector(self, rhs, out) -> None:
        """``Csr::mul_vector(&self, rhs:&[T], out:&mut [T]) -> Result<(),MatErr>``
        (sparse.rs:468-482): dense slice in, dense slice out (no zero-drop), computed by the same
        GPU kernels as a one-column ``mul_dense``."""
        rhs = np.ascontiguousarray(rhs, dtype=self.dtype)
        if self.dims.cols != rhs.shape[0] or self.dims.rows != out.shape[0]:
            raise MatError(MatErr.IncorrectDimensions)                   # sparse.rs:469-471
        self._check_multipliable()
        from .gpu import DeviceCsr
        with DeviceCsr.from_host(self) as a:
            res = np.zeros(self.dims.rows, dtype=self.dtype)
            sfx = _lib.suffix(self.dtype)
            _lib.check(getattr(_lib.lib(), f"bsm_mul_vector_{sfx}")(
                a.handle, _lib.ptr(rhs), rhs.shape[0], _lib.ptr(res), res.shape[0]))
        out[:] = res
