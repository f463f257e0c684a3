"""Host ``DenseS<T, ROWS, COLS>`` — the reference's fixed-size column-major operand
(/root/reference/src/dense_static.rs), the right-hand side of ``Csr::mul_dense_s`` (sparse.rs:448-466).

Rust's const generics become the two leading constructor arguments ``(ROWS, COLS)``; the storage is one
``(COLS, ROWS)`` C-contiguous array, i.e. ``[[T; ROWS]; COLS]``: ``data[c]`` is COLUMN ``c``.

The reference's ``from_data`` has a quirk this mirror keeps (dense_static.rs:21-35): ``col_count`` / ``row_count``
(what ``get_dims`` reports) are taken from the SLICES passed in, while only the ``ROWS x COLS`` window of them is copied;
slices shorter than the window are an index panic there and an ``IndexError`` here."""
from __future__ import annotations

import numpy as np

from .util import GetDims, MatDim


class DenseS(GetDims):
    def __init__(self, rows: int, cols: int, data: np.ndarray, col_count: int | None = None, row_count: int | None = None):
        self.ROWS = int(rows)
        self.COLS = int(cols)
        if data.shape != (self.COLS, self.ROWS):
            raise ValueError("DenseS storage is [[T; ROWS]; COLS]")
        self.data = data
        self.col_count = self.COLS if col_count is None else int(col_count)
        self.row_count = self.ROWS if row_count is None else int(row_count)

    # dense_static.rs:13-15
    @classmethod
    def new_default(cls, rows: int, cols: int, dtype=np.float64) -> "DenseS":
        return cls.new(np.dtype(dtype).type(0), rows, cols, dtype)

    # dense_static.rs:17-19
    @classmethod
    def new(cls, val, rows: int, cols: int, dtype=None) -> "DenseS":
        dt = np.dtype(dtype) if dtype is not None else np.asarray(val).dtype
        return cls(rows, cols, np.full((int(cols), int(rows)), val, dtype=dt))

    # dense_static.rs:21-35 — data[c] is COLUMN c; dims come from the slices, the copy from the const parameters
    @classmethod
    def from_data(cls, data, rows: int | None = None, cols: int | None = None, dtype=np.float64) -> "DenseS":
        col_count = len(data)
        row_count = len(data[0])                                  # data[0] of an empty slice panics in the reference too
        rows = row_count if rows is None else int(rows)
        cols = col_count if cols is None else int(cols)
        temp = np.zeros((cols, rows), dtype=dtype)
        for i in range(cols):
            if i >= col_count or len(data[i]) < rows:
                raise IndexError("DenseS.from_data: the slices are smaller than ROWS x COLS")
            temp[i, :] = np.asarray(data[i][:rows], dtype=dtype)
        return cls(rows, cols, temp, col_count, row_count)

    @property
    def dtype(self):
        return self.data.dtype

    def get_col(self, col_index: int) -> np.ndarray:              # dense_static.rs:37-39
        if not 0 <= col_index < self.COLS:
            raise IndexError("DenseS.get_col: column out of range")
        return self.data[col_index]

    def get_col_mut(self, col_index: int) -> np.ndarray:          # dense_static.rs:41-43
        return self.get_col(col_index)

    def get_dims(self) -> MatDim:                                 # dense_static.rs:46-53
        return MatDim(rows=self.row_count, cols=self.col_count)

    def __eq__(self, other):                                      # #[derive(PartialEq)]  dense_static.rs:4
        return (isinstance(other, DenseS) and (self.ROWS, self.COLS) == (other.ROWS, other.COLS)
                and self.col_count == other.col_count and self.row_count == other.row_count
                and np.array_equal(self.data, other.data))

    def __str__(self):                                            # Display  dense_static.rs:55-68
        lines = []
        for r in range(self.row_count):
            lines.append("|" + "".join(f"{self.get_col(c)[r]:>5}" for c in range(self.col_count)) + "|")
        return "\n".join(lines) + ("\n" if lines else "")

    def __repr__(self):
        return f"DenseS<{self.dtype}, {self.ROWS}, {self.COLS}>"
