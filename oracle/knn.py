"""Exact-kNN oracle (test infrastructure only; see oracle/__init__.py).

`knn_brute` is the reference's exact path verbatim: scanpy builds
`KNeighborsTransformer(n_neighbors=min(n-1, k), algorithm='brute', metric='euclidean')`
(src/scanpy/neighbors/__init__.py:754-768; sklearn's transformer returns k+1 entries per row,
the query point itself included, which `_get_indices_distances_from_sparse_matrix` trims to k) and post-processes with `_get_indices_distances_from_sparse_matrix`
and `_get_sparse_matrix_from_indices_distances` (src/scanpy/neighbors/_common.py:35-98),
restated below.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse


def knn_brute(x: np.ndarray, n_neighbors: int, *, n_jobs: int = -1):
    """(indices[n,k], distances[n,k]) with self in column 0; k = n_neighbors incl. self."""
    from sklearn.neighbors import KNeighborsTransformer

    n = x.shape[0]
    t = KNeighborsTransformer(n_neighbors=min(n - 1, n_neighbors), algorithm="brute",
                              metric="euclidean", n_jobs=n_jobs)
    d = t.fit_transform(x)
    return indices_distances_from_sparse(d, n_neighbors)


def knn_brute_queries(x_all: np.ndarray, q0: int, q1: int, n_neighbors: int, *, n_jobs: int = -1):
    """Exact kNN (incl. self) of rows [q0,q1) against all rows of x_all (for slab timing)."""
    from sklearn.neighbors import NearestNeighbors

    nn = NearestNeighbors(n_neighbors=n_neighbors, algorithm="brute", metric="euclidean", n_jobs=n_jobs)
    nn.fit(x_all)
    dist, idx = nn.kneighbors(x_all[q0:q1])
    return idx, dist


def indices_distances_from_sparse(d, n_neighbors: int):
    """src/scanpy/neighbors/_common.py:74-98 + the shortcut :126-143 (constant nnz/row only)."""
    nnzs = np.diff(d.indptr)
    assert (nnzs == nnzs[0]).all(), "oracle handles constant-nnz rows only"
    n_obs, k = d.shape[0], int(nnzs[0])
    indices = d.indices.reshape(n_obs, k)
    distances = d.data.reshape(n_obs, k)
    if not (indices[:, 0] == np.arange(n_obs)).any():  # _has_self_column, :17-22
        indices = np.hstack([np.arange(n_obs)[:, None], indices])
        distances = np.hstack([np.zeros(n_obs)[:, None], distances])
    if indices.shape[1] > n_neighbors:
        indices, distances = indices[:, :n_neighbors], distances[:, :n_neighbors]
    return indices, distances


def sparse_from_indices_distances(indices, distances, *, keep_self: bool = False):
    """src/scanpy/neighbors/_common.py:35-61."""
    if not keep_self:
        assert (indices[:, 0] == np.arange(indices.shape[0])).any()
        indices, distances = indices[:, 1:], distances[:, 1:]
    indptr = np.arange(0, indices.size + 1, indices.shape[1])
    return sparse.csr_matrix((distances.copy().ravel(), indices.copy().ravel(), indptr),
                             shape=(indices.shape[0],) * 2)


def same_neighbor_sets(idx_a, dist_a, idx_b, dist_b, *, rtol: float = 1e-6) -> np.ndarray:
    """Per-row bool: neighbour SETS equal, where members whose distance ties (rel. rtol) with the
    k-th distance may be swapped (argpartition order among ties is unspecified in sklearn)."""
    n, k = idx_a.shape
    ok = np.zeros(n, bool)
    for i in range(n):
        sa, sb = set(idx_a[i].tolist()), set(idx_b[i].tolist())
        if sa == sb:
            ok[i] = True
            continue
        kth = max(dist_a[i].max(), dist_b[i].max())
        da = {j: d for j, d in zip(idx_a[i], dist_a[i])}
        db = {j: d for j, d in zip(idx_b[i], dist_b[i])}
        diff = (sa - sb) | (sb - sa)
        ok[i] = all(abs((da.get(j, db.get(j))) - kth) <= rtol * max(kth, 1e-30) for j in diff)
    return ok
