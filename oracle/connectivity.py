"""`method='gauss'` / `method='jaccard'` connectivities oracle (test infrastructure only).

numpy restatements of src/scanpy/neighbors/_connectivity.py:17-100 (gauss, the sparse kNN branch that
`sc.pp.neighbors(method='gauss', knn=True)` takes) and :141-186 (jaccard).  Pinned by the reference's 4-point goldens
`connectivities_gauss_knn` / `connectivities_jaccard` (tests/test_neighbors.py:66-72,120-126,195-226).
"""
from __future__ import annotations

import numpy as np
from scipy import sparse


def gauss_knn(knn_indices, knn_dists):
    """knn_indices/knn_dists [n,k] with self in column 0 -> CSR of gaussian weights, 'copy if missing' symmetrised."""
    n, k = knn_indices.shape
    idx = knn_indices[:, 1:]
    d_sq = np.asarray(knn_dists, np.float64)[:, 1:] ** 2
    sigmas_sq = np.median(d_sq, axis=1)
    sigmas = np.sqrt(sigmas_sq)
    num = 2 * sigmas[:, None] * sigmas[idx]
    den = sigmas_sq[:, None] + sigmas_sq[idx]
    w = np.sqrt(num / den) * np.exp(-d_sq / den)
    m = sparse.lil_matrix((n, n))
    for i in range(n):
        for j, v in zip(idx[i], w[i]):
            m[i, j] = v
    sets = [set(r.tolist()) for r in idx]
    for i in range(n):
        for j, v in zip(idx[i], w[i]):
            if i not in sets[j]:
                m[j, i] = v
    return m.tocsr()


def jaccard(knn_indices):
    n, k = knn_indices.shape
    adjacency = sparse.csr_matrix((np.ones(n * (k - 1)), knn_indices[:, 1:].ravel(), np.arange(0, n * (k - 1) + 1, k - 1)),
                                  shape=(n, n))
    i_idx = np.repeat(np.arange(n), k - 1)
    j_idx = knn_indices[:, 1:].ravel()
    shared = np.asarray(adjacency[i_idx, :].multiply(adjacency[j_idx, :]).sum(axis=1)).ravel()
    jac = shared / (2 * (k - 1) - shared)
    mask = jac != 0
    c = sparse.csr_matrix((jac[mask], (i_idx[mask], j_idx[mask])), shape=(n, n))
    return (c + c.T) / 2
