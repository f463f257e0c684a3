"""Host mirror of the substitution half of the reference's ``solve`` (src/lib.rs:11-65), running on the GPU.

``solve(a, b)`` in the reference is ``cholesky_decomp`` -> ``transpose`` -> ``forward_substitution`` ->
``backward_substitution``.  The factorisation and the transpose are out of scope of this package (SURVEY §2 rows 10, 12:
dense-style O(n^4) as coded); the two substitutions — what is left once the factor exists — run on the device through
``bsm_forward_substitution`` / ``bsm_backward_substitution``, bit-identical to the reference's loops.
"""
from __future__ import annotations

from .dense import Dense
from .sparse import Csr
from .util import MatErr, MatError


def _check(l: Csr, b: Dense):
    if not l.is_finalised:
        raise MatError(MatErr.MatrixNotFinalised, "finalise() the factor first")
    if l.dims.rows != l.dims.cols:
        raise MatError(MatErr.NonSquareMatrix)
    if b.get_dims().rows != l.dims.rows:
        raise MatError(MatErr.IncorrectDimensions)
    if b.dtype != l.dtype:
        raise TypeError("factor and right-hand side must have the same element type")


def forward_substitution(l: Csr, b: Dense) -> Dense:
    """``forward_substitution(l: Csr<f32>, b: Dense<f32>) -> Dense<f32>`` (lib.rs:28-46): solve L y = b."""
    from .gpu import DeviceCsr, DeviceDense
    _check(l, b)
    with DeviceCsr.from_host(l) as dl, DeviceDense.from_host(b) as db, dl.forward_substitution(db) as dy:
        return dy.to_host()


def backward_substitution(l_star: Csr, y: Dense) -> Dense:
    """``backward_substitution(l_star, y)`` (lib.rs:49-65): solve L* x = y."""
    from .gpu import DeviceCsr, DeviceDense
    _check(l_star, y)
    with DeviceCsr.from_host(l_star) as dl, DeviceDense.from_host(y) as dy, dl.backward_substitution(dy) as dx:
        return dx.to_host()


def solve_with_factor(l: Csr, l_star: Csr, b: Dense) -> Dense:
    """The tail of ``solve`` (lib.rs:22-23) once ``l = a.cholesky_decomp()`` and ``l_star = l.transpose()`` exist:
    both substitutions on the device, the intermediate ``y`` never leaves HBM."""
    from .gpu import DeviceCsr, DeviceDense
    _check(l, b)
    _check(l_star, b)
    with DeviceCsr.from_host(l) as dl, DeviceCsr.from_host(l_star) as dls, DeviceDense.from_host(b) as db:
        with dl.forward_substitution(db) as dy, dls.backward_substitution(dy) as dx:
            return dx.to_host()
