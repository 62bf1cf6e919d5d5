"""PCA oracle = the reference's arithmetic (test infrastructure only; see oracle/__init__.py).

scanpy's `pca()` for a scipy CSR input with zero_center=True and svd_solver in {None,'arpack'}
does exactly this (src/scanpy/preprocessing/_pca/__init__.py:282-291,308,338-363):

    pca_ = sklearn.decomposition.PCA(n_components=n_comps, svd_solver='arpack', random_state=0)
    x_pca = pca_.fit_transform(x);  x_pca = x_pca.astype('float32')

and reads `components_`, `explained_variance_`, `explained_variance_ratio_`.
`covariance_eigh` restates src/scanpy/preprocessing/_pca/_dask.py:24-132,143-213
(Gram -> cov with bias -> eigh -> project), in float64.
"""
from __future__ import annotations

import numpy as np


def pca_arpack(x, n_comps: int, *, random_state: int = 0, dtype="float32"):
    """Return dict(X_pca, components, variance, variance_ratio, mean) as scanpy would store them."""
    from sklearn.decomposition import PCA

    pca_ = PCA(n_components=n_comps, svd_solver="arpack", random_state=random_state)
    x_pca = pca_.fit_transform(x)
    if x_pca.dtype != np.dtype(dtype):
        x_pca = x_pca.astype(dtype)
    return dict(
        X_pca=np.ascontiguousarray(x_pca),
        components=pca_.components_,
        variance=pca_.explained_variance_,
        variance_ratio=pca_.explained_variance_ratio_,
        mean=np.asarray(pca_.mean_).ravel(),
        singular_values=pca_.singular_values_,
    )


def pca_exact_f64(x, n_comps: int):
    """Dense float64 PCA by eigh of the covariance: the 'true' answer both the reference's
    float32 ARPACK run and our CUDA path approximate (used to calibrate tolerances)."""
    import scipy.linalg as sla
    from scipy import sparse

    xd = x.toarray().astype(np.float64) if sparse.issparse(x) else np.asarray(x, np.float64)
    n = xd.shape[0]
    mu = xd.mean(axis=0)
    xc = xd - mu
    cov = xc.T @ xc / (n - 1)
    g = cov.shape[0]
    w, v = sla.eigh(cov, subset_by_index=[g - n_comps, g - 1])
    w, v = w[::-1], v[:, ::-1]
    vt = v.T.copy()
    # svd_flip(u_based_decision=False): largest-|.| entry of each row of Vt positive
    # (site-packages/sklearn/utils/extmath.py:974-981)
    sign = np.sign(vt[np.arange(n_comps), np.argmax(np.abs(vt), axis=1)])
    vt *= sign[:, None]
    total_var = xc.var(axis=0, ddof=1).sum()
    return dict(X_pca=xc @ vt.T, components=vt, variance=w, variance_ratio=w / total_var, mean=mu)


def align_signs(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Flip columns of `a` so that each correlates positively with the same column of `b`."""
    s = np.sign(np.einsum("ij,ij->j", a, b))
    s[s == 0] = 1
    return a * s


def gram_f64_blocked(x, *, block_rows: int = 16384):
    """(G = X^T X, column sums), both in float64, for a scipy CSR matrix, by dense row blocks through BLAS (a sparse-sparse product
    costs sum_r nnz_r^2 hashed updates on one core; densified syrk blocks use every core).  This is the host-side
    restatement of the reference's `csr_gram_dense` (src/scanpy/preprocessing/_pca/_kernels.py:14-58)."""
    n, g = x.shape
    gram = np.zeros((g, g), np.float64)
    colsum = np.zeros(g, np.float64)
    for r0 in range(0, n, block_rows):
        blk = x[r0:r0 + block_rows].toarray().astype(np.float64, copy=False)
        gram += blk.T @ blk
        colsum += blk.sum(axis=0)
    return gram, colsum


def pca_gram_f64(x, n_comps: int, *, rows=None):
    """Float64 ground truth through the covariance_eigh route (src/scanpy/preprocessing/_pca/_dask.py:143-213
    restated with ddof=1 like sklearn's explained_variance_): Gram -> covariance -> LAPACK eigh -> projection.
    Never densifies more than a row block, so it also serves the 100k and 1.3M configurations.
    `rows`: project only these rows (X_pca of a sample); None = all rows."""
    import scipy.linalg as sla

    n, g = x.shape
    gram, colsum = gram_f64_blocked(x)
    mu = colsum / n
    cov = (gram - n * np.outer(mu, mu)) / (n - 1)
    w, v = sla.eigh(cov, subset_by_index=[g - n_comps, g - 1])
    w, v = w[::-1], v[:, ::-1]
    vt = v.T.copy()
    sign = np.sign(vt[np.arange(n_comps), np.argmax(np.abs(vt), axis=1)])
    vt *= sign[:, None]
    xs = x if rows is None else x[np.asarray(rows)]
    x_pca = np.asarray(xs.astype(np.float64) @ vt.T) - mu @ vt.T
    total_var = float(np.trace(cov))
    # relative spectrum gaps (sigma_j - sigma_{j+1}) / sigma_j of the singular values, one more than n_comps
    w_all = sla.eigh(cov, subset_by_index=[g - n_comps - 1, g - 1], eigvals_only=True)[::-1]
    s = np.sqrt(np.maximum(w_all, 0.0))
    gaps = (s[:-1] - s[1:]) / s[:-1]
    return dict(X_pca=x_pca, components=vt, variance=w, variance_ratio=w / total_var, mean=mu, gaps=gaps)


def truncated_svd_arpack(x, n_comps: int, dtype="float64"):
    """The reference's `zero_center=False` call (src/scanpy/preprocessing/_pca/__init__.py:331-336) with the exact solver:
    sklearn TruncatedSVD(algorithm='arpack') -> dict(X_pca, components, variance, variance_ratio, singular_values)."""
    from sklearn.decomposition import TruncatedSVD

    t = TruncatedSVD(n_components=n_comps, algorithm="arpack", random_state=0)
    xp = t.fit_transform(x.astype(dtype))
    return dict(X_pca=xp, components=t.components_, variance=t.explained_variance_, variance_ratio=t.explained_variance_ratio_,
                singular_values=t.singular_values_)
