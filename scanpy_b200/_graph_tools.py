"""`sc.tl.umap`, `sc.tl.diffmap`, `sc.tl.paga` with scanpy's signatures on libscanpy_b200.

References: src/scanpy/tools/_umap.py:33-229, src/scanpy/tools/_diffmap.py:26-142,
src/scanpy/neighbors/__init__.py:791-884 (compute_transitions / compute_eigen), src/scanpy/tools/_paga.py:21-264.
The per-arc / per-vertex arithmetic (UMAP epochs, spectral initialisation, Lanczos eigensolver, PAGA arc counts) runs in
csrc/umap.cu, csrc/eigs.cu, csrc/graph.cu; what stays on the host is parameter-sized: the (a, b) curve fit of
`find_ab_params` (300 points) and PAGA's n_groups x n_groups statistics + spanning tree.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse

from . import _abi, _ops
from ._abi import check, ptr
from ._compat import LegacyRng, accepts_legacy_random_state, log_done, log_start, meta_random_state, seed_from_rng, warn


def find_ab_params(spread: float, min_dist: float) -> tuple[float, float]:
    """umap.umap_.find_ab_params (umap-learn, third party): fit 1 / (1 + a x^(2b)) to the fuzzy-set membership curve."""
    from scipy.optimize import curve_fit

    def curve(x, a, b):
        return 1.0 / (1.0 + a * x ** (2 * b))

    xv = np.linspace(0, spread * 3, 300)
    yv = np.zeros(xv.shape)
    yv[xv < min_dist] = 1.0
    yv[xv >= min_dist] = np.exp(-(xv[xv >= min_dist] - min_dist) / spread)
    params, _ = curve_fit(curve, xv, yv)
    return float(params[0]), float(params[1])


def _neighbors_view(adata, neighbors_key: str):
    nb = adata.uns[neighbors_key]
    conn_key = nb.get("connectivities_key", "connectivities")
    dist_key = nb.get("distances_key", "distances")
    return nb, conn_key, dist_key


@accepts_legacy_random_state(0)
def umap(adata, *, min_dist: float = 0.5, spread: float = 1.0, n_components: int = 2, maxiter: int | None = None,
         alpha: float = 1.0, gamma: float = 1.0, negative_sample_rate: int = 5, init_pos="spectral", rng=None,
         a: float | None = None, b: float | None = None, method: str = "umap", key_added: str | None = None,
         neighbors_key: str = "neighbors", copy: bool = False):
    """Embed the neighborhood graph using UMAP (signature of `scanpy.tl.umap`, tools/_umap.py:33-51)."""
    adata = adata.copy() if copy else adata
    key_uns, key_obsm = (key_added or "umap"), (key_added or "X_umap")  # _embedding_keys("umap", key_added)
    if neighbors_key is None:
        neighbors_key = "neighbors"
    if neighbors_key not in adata.uns:
        raise ValueError(f"Did not find .uns[{neighbors_key!r}]. Run `sc.pp.neighbors` first.")
    if method != "umap":
        if method == "rapids":
            raise NotImplementedError("method='rapids' is not available in scanpy_b200 (the device path IS method='umap')")
        raise ValueError(f"Unknown method {method}")
    start = log_start("computing UMAP")
    nb, conn_key, _ = _neighbors_view(adata, neighbors_key)
    if "params" not in nb or nb["params"].get("method") != "umap":
        warn(f'.obsp["{conn_key}"] have not been computed using umap')
    if a is None or b is None:
        a, b = find_ab_params(spread, min_dist)
    adata.uns[key_uns] = dict(params=dict(a=a, b=b, **meta_random_state(rng)))
    conn = adata.obsp[conn_key]
    n = conn.shape[0]
    if isinstance(init_pos, str) and init_pos in adata.obsm:
        init = adata.obsm[init_pos]
    elif isinstance(init_pos, str) and init_pos == "paga":
        raise NotImplementedError("init_pos='paga' needs `sc.pl.paga` layout positions, which scanpy_b200 does not compute")
    else:
        init = init_pos
    seed = seed_from_rng(rng)
    if isinstance(init, str):
        if init == "random":
            init = np.random.default_rng(seed).uniform(-10.0, 10.0, size=(n, n_components)).astype(np.float32)
        elif init != "spectral":
            raise ValueError(f"init_pos={init!r}: expected 'spectral', 'random', 'paga', an `.obsm` key or an array")
    else:
        init = np.asarray(init)
        if init.ndim != 2 or init.shape != (n, n_components):
            raise ValueError(f"init_pos must have shape {(n, n_components)}, got {init.shape}")
        init = np.ascontiguousarray(init, dtype=np.float32)  # check_array(init_coords, dtype=np.float32) (_umap.py:177-178)
    default_epochs = 500 if n <= 10000 else 200
    n_epochs = default_epochs if maxiter is None else int(maxiter)
    adj = conn.tocsr()
    if adj.dtype != np.float32:
        adj = adj.astype(np.float32)
    x_umap = _ops.umap_layout(adj, n_components=n_components, n_epochs=n_epochs, a=a, b=b, gamma=gamma, initial_alpha=alpha,
                              negative_sample_rate=negative_sample_rate, seed=seed, init=init)
    adata.obsm[key_obsm] = x_umap
    log_done(start, f"added\n    {key_obsm!r}, UMAP coordinates (adata.obsm)\n    {key_uns!r}, UMAP parameters (adata.uns)")
    return adata if copy else None


def _transition_scale_device(ctx, d_indptr, d_indices, d_w, n: int, density_normalize: bool = True):
    """s with T_sym = diag(s) W diag(s) (Neighbors.compute_transitions, neighbors/__init__.py:791-830)."""
    import torch

    d_s = torch.empty(n, dtype=torch.float64, device="cuda")
    check(ctx.lib.sb2_transition_scale_f64(ctx.handle, n, ptr(d_indptr), ptr(d_indices), ptr(d_w), int(density_normalize),
                                           ptr(d_s)))
    return d_s


@accepts_legacy_random_state(0)
def diffmap(adata, n_comps: int = 15, *, neighbors_key: str | None = None, key_added: str | None = None, rng=None,
            copy: bool = False):
    """Diffusion maps (signature of `scanpy.tl.diffmap`, tools/_diffmap.py:26-111)."""
    if neighbors_key is None:
        neighbors_key = "neighbors"
    if neighbors_key not in adata.uns:
        raise ValueError("You need to run `pp.neighbors` first to compute a neighborhood graph.")
    if n_comps <= 2:
        raise ValueError("Provide any value greater than 2 for `n_comps`. ")
    adata = adata.copy() if copy else adata
    key_uns, key_obsm = ("diffmap_evals", "X_diffmap") if key_added is None else (key_added, key_added)
    start = log_start(f"computing Diffusion Maps using {n_comps=}(=n_dcs)")
    _, conn_key, _ = _neighbors_view(adata, neighbors_key)
    conn = adata.obsp[conn_key].tocsr()
    if conn.dtype != np.float32:
        conn = conn.astype(np.float32)
    n = conn.shape[0]
    ctx = _abi.default_context()
    d_indptr, d_indices, d_w = _ops.csr_to_device(conn)
    d_s = _transition_scale_device(ctx, d_indptr, d_indices, d_w, n)
    n_comps = min(n - 1, n_comps)
    gen = rng.generator() if isinstance(rng, LegacyRng) else np.random.default_rng(rng)
    v0 = gen.standard_normal(n)  # compute_eigen: v0 = rng.standard_normal(n) (:868)
    evals, evecs, info = _ops.eigsh_scaled_device(ctx, d_indptr, d_indices, d_w, n, n_comps, d_scale=d_s, which="LM", v0=v0)
    if info["n_converged"] < n_comps:
        warn(f"diffmap: {n_comps - info['n_converged']} eigenpairs did not reach the residual tolerance "
             f"(max residual {info['max_residual']:.2e})")
    basis = _ops._to_host(evecs.t().contiguous().to(dtype=_ops._torch().float32))
    evals = evals.astype(np.float32)[::-1]  # sort='decrease' (:876-878)
    basis = basis[:, ::-1]
    adata.obsm[key_obsm] = np.ascontiguousarray(basis)
    adata.uns[key_uns] = evals if key_added is None else dict(evals=evals)
    log_done(start, f"added\n    {key_obsm!r}, diffmap coordinates (adata.obsm)\n    {key_uns!r}, eigenvalues of transition matrix (adata.uns)")
    return adata if copy else None


def paga(adata, groups: str | None = None, *, use_rna_velocity: bool = False, model: str = "v1.2",
         neighbors_key: str | None = None, copy: bool = False):
    """Partition-based graph abstraction (signature of `scanpy.tl.paga`, tools/_paga.py:21-157; model 'v1.2')."""
    import torch
    from scipy.sparse.csgraph import minimum_spanning_tree

    key = "neighbors" if neighbors_key is None else neighbors_key
    if key not in adata.uns:
        raise ValueError("You need to run `pp.neighbors` first to compute a neighborhood graph.")
    if groups is None:
        for k in ("leiden", "louvain"):
            if k in adata.obs.columns:
                groups = k
                break
    if groups is None:
        raise ValueError("You need to run `tl.leiden` or `tl.louvain` to compute community labels, or specify "
                         "`groups='an_existing_key'`")
    if groups not in adata.obs.columns:
        raise KeyError(f"`groups` key {groups!r} not found in `adata.obs`.")
    if use_rna_velocity:
        raise NotImplementedError("`use_rna_velocity=True` is not implemented in scanpy_b200.tl.paga")
    if model != "v1.2":
        if model == "v1.0":
            raise NotImplementedError("model='v1.0' is not implemented in scanpy_b200.tl.paga")
        raise ValueError(f"`model` {model} needs to be one of ('v1.2', 'v1.0').")
    adata = adata.copy() if copy else adata
    start = log_start("running PAGA")
    _, _, dist_key = _neighbors_view(adata, key)
    dist = adata.obsp[dist_key].tocsr()
    col = adata.obs[groups]
    if not hasattr(col, "cat"):
        col = col.astype("category")
    codes = np.ascontiguousarray(col.cat.codes.to_numpy(), dtype=np.int32)
    G = int(len(col.cat.categories))
    n = dist.shape[0]
    ctx = _abi.default_context()
    d_indptr = _ops._to_device(np.asarray(dist.indptr, np.int64))
    d_indices = _ops._to_device(np.asarray(dist.indices, np.int32))
    d_codes = _ops._to_device(codes)
    d_counts = torch.empty((G, G), dtype=torch.int64, device="cuda")
    check(ctx.lib.sb2_group_arc_counts(ctx.handle, n, ptr(d_indptr), ptr(d_indices), ptr(d_codes), G, ptr(d_counts)))
    counts = _ops._to_host(d_counts).astype(np.float64)
    # --- n_groups-sized statistics of _compute_connectivities_v1_2 (:188-208) ---
    ns = np.bincount(codes[codes >= 0], minlength=G)
    n_tot = int(ns.sum())
    es_inner = np.diag(counts).copy()
    inter = counts.copy()
    np.fill_diagonal(inter, 0.0)
    es = es_inner + inter.sum(axis=1)
    inter = inter + inter.T
    ii, jj = np.nonzero(inter)
    expected = (es[ii] * ns[jj] + es[jj] * ns[ii]) / (n_tot - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        scaled = np.where(expected != 0, inter[ii, jj] / expected, 1.0)
    scaled = np.minimum(scaled, 1.0)
    connectivities = sparse.csr_matrix((scaled, (ii, jj)), shape=(G, G))
    # _get_connectivities_tree_v1_2 (:236-250)
    inv = connectivities.copy()
    inv.data = 1.0 / inv.data
    mst = minimum_spanning_tree(inv).tocsr()
    tree = sparse.lil_matrix((G, G), dtype=float)
    for i in range(G):
        nbrs = mst[i].nonzero()[1]
        if len(nbrs) > 0:
            tree[i, nbrs] = connectivities[i, nbrs]
    if "paga" not in adata.uns:
        adata.uns["paga"] = {}
    adata.uns["paga"]["connectivities"] = connectivities
    adata.uns["paga"]["connectivities_tree"] = tree.tocsr()
    adata.uns[f"{groups}_sizes"] = np.array(ns)
    adata.uns["paga"]["groups"] = groups
    log_done(start, "added\n    'paga/connectivities', connectivities adjacency (adata.uns)\n"
                    "    'paga/connectivities_tree', connectivities subtree (adata.uns)")
    return adata if copy else None
