"""Host ``Dense<T>`` — the reference's column-major dense operand (/root/reference/src/dense.rs).

``data[c][r]``: a list of columns, each a contiguous 1-D numpy array (the ``Vec<Vec<T>>``).
Constructor argument order is ``(col_count, row_count)`` as in the reference (dense.rs:13,17) and
``from_data`` takes a sequence of COLUMNS (dense.rs:21-29)."""
from __future__ import annotations

import numpy as np

from .util import GetDims, MatDim


class Dense(GetDims):
    def __init__(self, col_count: int, row_count: int, data):
        self.col_count = int(col_count)
        self.row_count = int(row_count)
        self.data = data

    # dense.rs:13-15
    @classmethod
    def new_default_with_dims(cls, col_count: int, row_count: int, dtype=np.float64) -> "Dense":
        return cls.new_with_dims(np.dtype(dtype).type(0), col_count, row_count, dtype)

    # dense.rs:17-19
    @classmethod
    def new_with_dims(cls, val, col_count: int, row_count: int, dtype=None) -> "Dense":
        dt = np.dtype(dtype) if dtype is not None else np.asarray(val).dtype
        return cls(col_count, row_count, [np.full(row_count, val, dtype=dt) for _ in range(col_count)])

    # dense.rs:21-29 — data[c] is COLUMN c
    @classmethod
    def from_data(cls, data, dtype=np.float64) -> "Dense":
        cols = [np.array(c, dtype=dtype, copy=True).reshape(-1) for c in data]
        row_count = cols[0].shape[0]
        if any(c.shape[0] != row_count for c in cols):
            raise ValueError("all columns must have the same length")
        return cls(len(cols), row_count, cols)

    @classmethod
    def from_columns_nocopy(cls, columns) -> "Dense":
        """Adopt existing contiguous column arrays (e.g. pinned buffers) without copying."""
        cols = list(columns)
        return cls(len(cols), cols[0].shape[0] if cols else 0, cols)

    @property
    def dtype(self):
        return self.data[0].dtype if self.data else np.dtype(np.float64)

    def get_col(self, col_index: int) -> np.ndarray:          # dense.rs:31-33
        return self.data[col_index]

    def get_col_mut(self, col_index: int) -> np.ndarray:      # dense.rs:35-37 (numpy views are mutable)
        return self.data[col_index]

    def get_dims(self) -> MatDim:                             # dense.rs:40-47
        return MatDim(rows=self.row_count, cols=self.col_count)

    def to_rowmajor(self) -> np.ndarray:
        if not self.data:
            return np.zeros((self.row_count, 0))
        return np.stack(self.data, axis=1)

    def __eq__(self, other):                                  # #[derive(PartialEq)]  dense.rs:4
        return (isinstance(other, Dense) and self.col_count == other.col_count
                and self.row_count == other.row_count
                and all(np.array_equal(a, b) for a, b in zip(self.data, other.data)))

    def __str__(self):                                        # Display  dense.rs:49-62
        lines = []
        for r in range(self.row_count):
            lines.append("|" + "".join(f"{self.data[c][r]:>5}" for c in range(self.col_count)) + "|")
        return "\n".join(lines) + ("\n" if lines else "")

    def __repr__(self):
        return f"Dense(col_count={self.col_count}, row_count={self.row_count})"
