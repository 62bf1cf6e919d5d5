"""CPU oracles of the widened graph tools: UMAP layout, diffusion-map eigen-decomposition, PAGA connectivities.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never imported by scanpy_b200/.

* diffmap: the reference's own arithmetic, live — `Neighbors.compute_transitions` / `compute_eigen`
  (src/scanpy/neighbors/__init__.py:791-884) are numpy/scipy calls and are restated line by line around the same
  `scipy.sparse.linalg.eigsh` call (scipy is installed here).
* PAGA v1.2: `PAGA._compute_connectivities_v1_2` / `_get_connectivities_tree_v1_2` (src/scanpy/tools/_paga.py:177-250) with
  igraph's `VertexClustering.cluster_graph(combine_edges='sum')` replaced by a scipy group-indicator product.
* UMAP: umap-learn >= 0.5.12 (pyproject.toml:76) is a third-party dependency that is NOT vendored under /root/reference and
  not installed here; `simplicial_set_embedding` / `optimize_layout_euclidean` are restated from the published algorithm
  (McInnes et al. 2018; upstream umap/umap_.py, umap/layouts.py) as a SEQUENTIAL numba loop.  Coordinates cannot be pinned
  (the reference's tests hold no UMAP golden: tests/test_embedding.py:55-97 checks dtype-invariance, recorded params and
  that connectivities are untouched; results depend on seed and library version), QUALITY is: the reference's in-tree
  fixture stores `obsm/X_umap`, scanpy's own sc.tl.umap output on the stored connectivities, and this restatement matches
  it in trustworthiness (0.9513-0.9526 vs 0.9512), cluster separation (0.49-0.51 vs 0.506) and neighbourhood agreement
  (tests/test_oracle_goldens.py::test_umap_oracle_matches_the_reference_fixtures_stored_embedding).
"""
from __future__ import annotations

import numba
import numpy as np
from scipy import sparse
from scipy.sparse.linalg import eigsh


# ------------------------------------------------------------------------------------------ diffmap
def transitions_sym(conn, density_normalize: bool = True):
    """neighbors/__init__.py:791-830."""
    conn = sparse.csr_matrix(conn)
    if density_normalize:
        dens = np.asarray(conn.sum(axis=0))
        dens = sparse.spdiags(1.0 / dens, 0, conn.shape[0], conn.shape[0])
        conn_norm = dens @ conn @ dens
    else:
        conn_norm = conn
    z = np.sqrt(np.asarray(conn_norm.sum(axis=0)))
    Z = sparse.spdiags(1.0 / z, 0, conn_norm.shape[0], conn_norm.shape[0])
    return Z @ conn_norm @ Z


def diffmap_eigen(conn, n_comps: int = 15, seed: int = 0):
    """neighbors/__init__.py:832-884 with sort='decrease' -> (evals float32 [n_comps] decreasing, evecs float32 [n, n_comps])."""
    matrix = transitions_sym(conn)
    n_comps = min(matrix.shape[0] - 1, n_comps)
    matrix = matrix.astype(np.float64)
    rng = np.random.default_rng(seed)
    v0 = rng.standard_normal(matrix.shape[0])
    evals, evecs = eigsh(matrix, k=n_comps, which="LM", ncv=None, v0=v0)
    evals, evecs = evals.astype(np.float32), evecs.astype(np.float32)
    return evals[::-1], evecs[:, ::-1]


# ------------------------------------------------------------------------------------------ PAGA
def paga_v1_2(distances, codes):
    """-> (connectivities CSR [G,G], connectivities_tree CSR, group sizes); tools/_paga.py:177-250."""
    from scipy.sparse.csgraph import minimum_spanning_tree

    ones = sparse.csr_matrix(distances).copy()
    ones.data = np.ones(len(ones.data))
    codes = np.asarray(codes)
    G = int(codes.max()) + 1
    n = ones.shape[0]
    ind = sparse.csr_matrix((np.ones(n), (np.arange(n), codes)), shape=(n, G))
    cg = (ind.T @ ones @ ind).toarray()  # cg[a, b] = arcs from group a to group b (directed graph)
    ns = np.bincount(codes, minlength=G)
    es_inner = np.diag(cg).copy()
    inter = cg.copy()
    np.fill_diagonal(inter, 0)
    es = es_inner + inter.sum(axis=1)
    inter = inter + inter.T
    conn = np.zeros((G, G))
    for i, j in zip(*np.nonzero(inter)):
        exp = (es[i] * ns[j] + es[j] * ns[i]) / (n - 1)
        v = inter[i, j] / exp if exp != 0 else 1
        conn[i, j] = min(v, 1)
    conn = sparse.csr_matrix(conn)
    inv = conn.copy()
    inv.data = 1.0 / inv.data
    mst = minimum_spanning_tree(inv).tocsr()
    tree = sparse.lil_matrix((G, G), dtype=float)
    for i in range(G):
        nb = mst[i].nonzero()[1]
        if len(nb) > 0:
            tree[i, nb] = conn[i, nb]
    return conn, tree.tocsr(), ns


# ------------------------------------------------------------------------------------------ UMAP
def find_ab_params(spread: float, min_dist: float):
    from scipy.optimize import curve_fit

    def curve(x, a, b):
        return 1.0 / (1.0 + a * x ** (2 * b))

    xv = np.linspace(0, spread * 3, 300)
    yv = np.zeros(xv.shape)
    yv[xv < min_dist] = 1.0
    yv[xv >= min_dist] = np.exp(-(xv[xv >= min_dist] - min_dist) / spread)
    params, _ = curve_fit(curve, xv, yv)
    return params[0], params[1]


def spectral_layout(graph, dim: int):
    """umap/spectral.py `spectral_layout` (connected graph): eigenvectors 1..dim of the normalised Laplacian."""
    n = graph.shape[0]
    deg = np.asarray(graph.sum(axis=0)).ravel()
    D = sparse.spdiags(1.0 / np.sqrt(deg), 0, n, n)
    L = sparse.identity(n) - D @ graph @ D
    k = dim + 1
    vals, vecs = eigsh(L.astype(np.float64), k, which="SM", ncv=max(2 * k + 1, int(np.sqrt(n))), tol=1e-4, v0=np.ones(n),
                       maxiter=n * 5)
    order = np.argsort(vals)[1:k]
    return vecs[:, order]


@numba.njit(cache=False)
def _tau_rand_int(state):
    state[0] = (((state[0] & 4294967294) << 12) & 0xFFFFFFFF) ^ ((((state[0] << 13) & 0xFFFFFFFF) ^ state[0]) >> 19)
    state[1] = (((state[1] & 4294967288) << 4) & 0xFFFFFFFF) ^ ((((state[1] << 2) & 0xFFFFFFFF) ^ state[1]) >> 25)
    state[2] = (((state[2] & 4294967280) << 17) & 0xFFFFFFFF) ^ ((((state[2] << 3) & 0xFFFFFFFF) ^ state[2]) >> 11)
    return state[0] ^ state[1] ^ state[2]


@numba.njit(cache=False)
def _clip(v):
    if v > 4.0:
        return 4.0
    if v < -4.0:
        return -4.0
    return v


@numba.njit(cache=False)
def _optimize(emb, head, tail, n_epochs, n_vertices, eps, a, b, rng_state, gamma, initial_alpha, neg_rate):
    dim = emb.shape[1]
    alpha = initial_alpha
    eps_neg = eps / neg_rate
    eonns = eps_neg.copy()
    eons = eps.copy()
    for n in range(n_epochs):
        for i in range(eps.shape[0]):
            if eons[i] <= n:
                j = head[i]
                k = tail[i]
                d2 = 0.0
                for d in range(dim):
                    d2 += (emb[j, d] - emb[k, d]) ** 2
                gc = -2.0 * a * b * d2 ** (b - 1.0) / (a * d2**b + 1.0) if d2 > 0.0 else 0.0
                for d in range(dim):
                    g = _clip(gc * (emb[j, d] - emb[k, d]))
                    emb[j, d] += g * alpha
                    emb[k, d] += -g * alpha
                eons[i] += eps[i]
                n_neg = int((n - eonns[i]) / eps_neg[i])
                for _p in range(n_neg):
                    k = _tau_rand_int(rng_state) % n_vertices
                    d2 = 0.0
                    for d in range(dim):
                        d2 += (emb[j, d] - emb[k, d]) ** 2
                    if d2 > 0.0:
                        gc = 2.0 * gamma * b / ((0.001 + d2) * (a * d2**b + 1.0))
                    elif j == k:
                        continue
                    else:
                        gc = 0.0
                    for d in range(dim):
                        g = _clip(gc * (emb[j, d] - emb[k, d])) if gc > 0.0 else 0.0
                        emb[j, d] += g * alpha
                eonns[i] += n_neg * eps_neg[i]
        alpha = initial_alpha * (1.0 - (n + 1.0) / n_epochs)
    return emb


def simplicial_set_embedding(graph, *, n_components=2, n_epochs=None, a=None, b=None, gamma=1.0, initial_alpha=1.0,
                             negative_sample_rate=5, init="spectral", seed=0):
    """umap/umap_.py `simplicial_set_embedding` (euclidean output metric, no densmap), sequential optimiser."""
    graph = sparse.coo_matrix(graph).copy()
    graph.sum_duplicates()
    n = graph.shape[1]
    default_epochs = 500 if n <= 10000 else 200
    if n_epochs is None:
        n_epochs = default_epochs
    thr = graph.data.max() / float(n_epochs if n_epochs > 10 else default_epochs)
    graph.data[graph.data < thr] = 0.0
    graph.eliminate_zeros()
    rs = np.random.RandomState(seed)
    if isinstance(init, str) and init == "spectral":
        initialisation = spectral_layout(sparse.csr_matrix(graph), n_components)
        expansion = 10.0 / np.abs(initialisation).max()
        emb = (initialisation * expansion).astype(np.float32) + rs.normal(scale=0.0001, size=[n, n_components]).astype(np.float32)
    elif isinstance(init, str) and init == "random":
        emb = rs.uniform(low=-10.0, high=10.0, size=(n, n_components)).astype(np.float32)
    else:
        emb = np.array(init, dtype=np.float32)
    eps = np.full(graph.data.shape[0], -1.0)
    n_samples = n_epochs * (graph.data / graph.data.max())
    eps[n_samples > 0] = float(n_epochs) / n_samples[n_samples > 0]
    rng_state = rs.randint(np.iinfo(np.int32).min + 1, np.iinfo(np.int32).max - 1, 3).astype(np.int64)
    emb = (10.0 * (emb - emb.min(0)) / (emb.max(0) - emb.min(0))).astype(np.float32, order="C")
    if a is None or b is None:
        a, b = find_ab_params(1.0, 0.5)
    return _optimize(emb, graph.row.astype(np.int64), graph.col.astype(np.int64), n_epochs, n, eps, a, b, rng_state, gamma,
                     initial_alpha, negative_sample_rate)
