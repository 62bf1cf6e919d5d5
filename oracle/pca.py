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
