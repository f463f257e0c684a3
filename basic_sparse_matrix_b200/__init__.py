"""basic_sparse_matrix_b200 — B200-native (sm_100a) Csr x Dense (SpMM / SpMV).

A from-scratch implementation of the one data-parallel hot path of
jamieapps101/Basic_Sparse_Matrix — ``Csr<T>::mul_dense`` (src/sparse.rs:426-446) — behind the
crate's own ``Csr`` / ``Dense`` / ``MatDim`` / ``MatErr`` API.  The product is the CUDA shared
library ``lib/libbsm_b200.so`` (C ABI in ``include/bsm.h``); this package is its host-side mirror
of the reference interface.  No CPU fallback exists.
"""
from .dense import Dense
from .dense_static import DenseS
from .sparse import Csr, CsrEntry
from .util import GetDims, MatDim, MatErr, MatError

__all__ = ["Csr", "CsrEntry", "Dense", "DenseS", "GetDims", "MatDim", "MatErr", "MatError", "gpu", "gen", "solve"]


def __getattr__(name):
    if name in ("gpu", "gen", "solve"):
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
