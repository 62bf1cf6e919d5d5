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


def knn_exact_f64(x_all: np.ndarray, rows, k: int, *, chunk: int = 256, slack: int = 16):
    """fp64 brute force for the query rows `rows` against ALL points of x_all: the set oracle that makes
    "identical kNN index sets" literal (sklearn's brute path evaluates |q|^2 - 2 q.c + |c|^2 on the input
    dtype and is itself inexact for float32 data).

    Candidates come from the GEMM form evaluated in float64 (k + slack smallest per row), then their squared
    distances are recomputed EXACTLY-rounded as sum_j (q_j - c_j)^2 in float64 and sorted by (d2, id).
    Returns (idx [m, k+1], d2 [m, k+1]): one column more than k, so a genuine tie at the k-th distance
    (d2[:, k-1] == d2[:, k], i.e. duplicated points) can be told from a wrong set."""
    x = np.ascontiguousarray(x_all, dtype=np.float64)
    rows = np.asarray(rows, dtype=np.int64)
    n = x.shape[0]
    kk = min(n, k + 1)
    take = min(n, k + 1 + slack)
    cn = np.einsum("ij,ij->i", x, x)
    out_i = np.empty((len(rows), kk), np.int64)
    out_d = np.empty((len(rows), kk), np.float64)
    for s in range(0, len(rows), chunk):
        r = rows[s:s + chunk]
        q = x[r]
        d2 = cn[None, :] - 2.0 * (q @ x.T)              # |q|^2 is constant per row: irrelevant for the order
        if take < n:
            cand = np.argpartition(d2, take - 1, axis=1)[:, :take]
        else:
            cand = np.tile(np.arange(n), (len(r), 1))
        diff = x[cand] - q[:, None, :]
        ex = np.einsum("ijk,ijk->ij", diff, diff)        # exact-form squared distances of the candidates
        order = np.lexsort((cand, ex), axis=1)           # by (d2, id) like the CUDA path
        cand = np.take_along_axis(cand, order, 1)
        ex = np.take_along_axis(ex, order, 1)
        # the GEMM form's error (~1e-13 |x|^2) could misorder candidates right at the edge of `take`; the slack
        # makes that irrelevant unless more than `slack` points tie with the (k+1)-th distance to ~1e-13
        out_i[s:s + chunk] = cand[:, :kk]
        out_d[s:s + chunk] = ex[:, :kk]
    return out_i, out_d


def exact_set_mismatches(idx_got: np.ndarray, oracle_idx: np.ndarray, oracle_d2: np.ndarray, k: int) -> np.ndarray:
    """Per-row bool "neighbour set differs from the exact fp64 one".  idx_got [m, k] (self included), oracle
    arrays from `knn_exact_f64` (k+1 columns).  A row only passes with a different set when the exact k-th and
    (k+1)-th squared distances are EQUAL (duplicated points: the set is genuinely not unique) and every member
    that differs sits exactly at that tied distance."""
    m = idx_got.shape[0]
    bad = np.zeros(m, bool)
    has_next = oracle_idx.shape[1] > k
    for i in range(m):
        sa, so = set(idx_got[i, :k].tolist()), set(oracle_idx[i, :k].tolist())
        if sa == so:
            continue
        if not has_next or oracle_d2[i, k - 1] != oracle_d2[i, k]:
            bad[i] = True
            continue
        tied = {int(j) for j, d in zip(oracle_idx[i], oracle_d2[i]) if d == oracle_d2[i, k - 1]}
        # members outside the oracle's k+1 window cannot be verified here -> count as mismatch
        bad[i] = not ((sa - so) <= tied and (so - sa) <= tied)
    return bad
