"""Dims / error vocabulary of the reference crate (/root/reference/src/util.rs)."""
from __future__ import annotations

import enum
from dataclasses import dataclass


@dataclass(frozen=True)
class MatDim:
    """``MatDim { rows, cols }`` (util.rs:11-15); ``From<(usize,usize)>`` is ``(rows, cols)``
    (util.rs:23-27)."""
    rows: int
    cols: int

    def transpose(self) -> "MatDim":          # util.rs:17-21
        return MatDim(self.cols, self.rows)

    @staticmethod
    def of(d) -> "MatDim":
        if isinstance(d, MatDim):
            return d
        r, c = d
        return MatDim(int(r), int(c))

    def __iter__(self):                         # From<MatDim> for (usize,usize)  util.rs:29-33
        yield self.rows
        yield self.cols

    def __str__(self):                          # Display  util.rs:35-41
        return f"(rows: {self.rows}, cols: {self.cols})"


class MatErr(enum.Enum):
    """``enum MatErr`` (util.rs:47-55) — variants unchanged."""
    MatrixFinalised = 0
    MatrixNotFinalised = 1
    NonSquareMatrix = 2
    IncorrectDimensions = 3
    PaddingSizeSmallerThanOriginal = 4
    OutOfBounds = 5


class MatError(Exception):
    """Python spelling of ``Err(MatErr::X)``: ``e.kind`` is the MatErr variant."""

    def __init__(self, kind: MatErr, message: str = ""):
        super().__init__(f"{kind.name}{': ' + message if message else ''}")
        self.kind = kind

    def __eq__(self, other):
        return isinstance(other, MatError) and other.kind == self.kind

    def __hash__(self):
        return hash(self.kind)


class GetDims:
    """``trait GetDims`` (util.rs:43-45)."""

    def get_dims(self) -> MatDim:  # pragma: no cover - interface
        raise NotImplementedError
