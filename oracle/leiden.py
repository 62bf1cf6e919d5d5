"""ctypes wrapper around oracle/leiden_ref.c (test infrastructure only; see oracle/__init__.py).

Mirrors what scanpy's `leiden()` extracts from the back-end: `.membership` and `.modularity`
(src/scanpy/tools/_leiden.py:198,219).  Label level: the reference's tests hold no Leiden golden, but its in-tree fixture
pbmc68k_reduced stores `obs/louvain` (scanpy's own sc.tl.louvain output on the stored connectivities); on the unweighted
graph this oracle reproduces it at ARI 0.94-0.98 (tests/test_oracle_leiden_guarantees.py).
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_lib = None


def build() -> Path:
    so = _HERE / "libleiden_ref.so"
    src = _HERE / "leiden_ref.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "libleiden_ref.so"], check=True, capture_output=True)
    return so


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(str(build()))
        P = ctypes.c_void_p
        lib.leiden_ref.restype = ctypes.c_int
        lib.leiden_ref.argtypes = [ctypes.c_int32, P, P, P, ctypes.c_double, ctypes.c_int32, ctypes.c_uint64,
                                   P, P, P, P]
        lib.leiden_ref2.restype = ctypes.c_int
        lib.leiden_ref2.argtypes = [ctypes.c_int32, P, P, P, ctypes.c_double, ctypes.c_int32, ctypes.c_uint64, ctypes.c_double,
                                    P, P, P, P]
        lib.modularity_ref.restype = ctypes.c_double
        lib.modularity_ref.argtypes = [ctypes.c_int32, P, P, P, ctypes.c_double, P]
        _lib = lib
    return _lib


def _csr_args(adj):
    adj = adj.tocsr()
    indptr = np.ascontiguousarray(adj.indptr, np.int64)
    indices = np.ascontiguousarray(adj.indices, np.int32)
    w = np.ascontiguousarray(adj.data, np.float64)
    return adj.shape[0], indptr, indices, w


def leiden(adj, *, resolution: float = 1.0, n_iterations: int = -1, seed: int = 0, beta: float = 0.0):
    """-> (membership int32[n] renumbered by decreasing size, modularity at `resolution`, n_passes).

    beta = 0: greedy refinement, what leidenalg's optimiser does for scanpy's flavor='leidenalg';
    beta > 0: the paper's randomised refinement over well-connected vertices/communities, what igraph's
    community_leiden does for flavor='igraph' (scanpy passes igraph's default beta = 0.01)."""
    lib = _load()
    n, indptr, indices, w = _csr_args(adj)
    member = np.empty(n, np.int32)
    q = ctypes.c_double()
    nc = ctypes.c_int32()
    passes = ctypes.c_int32()
    rc = lib.leiden_ref2(n, indptr.ctypes.data, indices.ctypes.data, w.ctypes.data, float(resolution),
                         int(n_iterations), int(seed), float(beta), member.ctypes.data, ctypes.addressof(q),
                         ctypes.addressof(nc), ctypes.addressof(passes))
    assert rc == 0
    return member, q.value, passes.value


def modularity(adj, membership, *, resolution: float = 1.0) -> float:
    lib = _load()
    n, indptr, indices, w = _csr_args(adj)
    m = np.ascontiguousarray(membership, np.int32)
    return lib.modularity_ref(n, indptr.ctypes.data, indices.ctypes.data, w.ctypes.data, float(resolution),
                              m.ctypes.data)
