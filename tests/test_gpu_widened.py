"""GPU tests of the widened rows: pp.scale, tl.louvain, metrics.modularity, tl.umap, tl.diffmap, tl.paga,
chunked PCA.  Same rule as test_gpu_parity.py: every product call goes through the C ABI, the oracle is the checker."""
import warnings

import numpy as np
import pandas as pd
import pytest
from scipy import sparse
from sklearn.metrics import adjusted_rand_score

import scanpy_b200 as sb
from oracle import leiden as old, preprocess as opre
from scanpy_b200._synth import synth_scipy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scaling_literals():
    from conftest import GOLDEN

    return np.load(GOLDEN / "reference_scaling_literals.npz")


@pytest.fixture(scope="module")
def graph_small():
    x, lab = synth_scipy(5000, 600, n_clusters=10, r=40)
    ad = sb.MiniAnnData(x)
    sb.pp.pca(ad, n_comps=30)
    sb.pp.neighbors(ad, n_neighbors=15)
    return ad, lab


# ------------------------------------------------------------------------------------------ pp.scale
@pytest.mark.parametrize("typ", [np.array, sparse.csr_matrix, sparse.csc_matrix], ids=lambda t: t.__name__)
@pytest.mark.parametrize("container", ["anndata", "array"])
@pytest.mark.parametrize("dtype", [np.float32, np.int64])
@pytest.mark.parametrize("zero_center", [True, False], ids=["center", "no_center"])
@pytest.mark.parametrize("masked", [False, True], ids=["no_mask", "mask"])
def test_scale_reference_goldens(scaling_literals, typ, container, dtype, zero_center, masked):
    # the reference's own matrix: tests/test_scaling.py:72-116
    L = scaling_literals
    x0 = L["X_for_mask"] if masked else L["X_original"]
    mask = np.array((0, 0, 1, 1, 1, 0, 0), dtype=bool) if masked else None
    expected = (L["X_centered_for_mask"] if zero_center else L["X_scaled_for_mask"]) if masked else (
        L["X_centered_original"] if zero_center else L["X_scaled_original"])
    x = typ(x0.astype(dtype))
    data = sb.MiniAnnData(x) if container == "anndata" else x
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = sb.pp.scale(data, zero_center=zero_center, copy=container == "array", mask_obs=mask)
    got = data.X if out is None else out
    got = got.toarray() if sparse.issparse(got) else np.asarray(got)
    assert np.allclose(got, expected)


def test_scale_clip_goldens_and_mask_string(scaling_literals):
    L = scaling_literals
    # tests/test_scaling.py:146-175 (clipped goldens) and :119-126 (string mask)
    out = sb.pp.scale(sparse.csr_matrix(L["X_original"].astype(np.float32)), zero_center=False, max_value=1)
    assert np.allclose(out.toarray(), L["X_scaled_original_clipped"])
    out = sb.pp.scale(L["X_for_mask"].astype(np.float32), zero_center=False, max_value=1,
                      mask_obs=np.array((0, 0, 1, 1, 1, 0, 0), dtype=bool))
    assert np.allclose(out, L["X_scaled_for_mask_clipped"])
    with pytest.raises(ValueError, match=r"Cannot.*refer.*mask.*without.*anndata"):
        sb.pp.scale(L["X_original"], mask_obs="mask")
    ad = sb.MiniAnnData(L["X_for_mask"].astype(np.float32))
    ad.obs["some cells"] = np.array((0, 0, 1, 1, 1, 0, 0), dtype=bool)
    sb.pp.scale(ad, mask_obs="some cells")
    assert np.array_equal(ad.X, L["X_centered_for_mask"])
    assert "mean of some cells" in ad.var.columns


@pytest.mark.parametrize("zero_center", [True, False])
@pytest.mark.parametrize("max_value", [None, 2.5])
def test_scale_matches_oracle_on_synthetic(zero_center, max_value):
    x, _ = synth_scipy(3000, 500, n_clusters=6, r=32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = sb.pp.scale(x, zero_center=zero_center, max_value=max_value, copy=True)
    ref, mean, std = opre.scale(x, zero_center=zero_center, max_value=max_value)
    if zero_center:
        assert got.dtype == np.float64 and not sparse.issparse(got)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)
        if max_value is not None:
            assert got.min() >= -max_value and got.max() <= max_value
    else:
        assert sparse.issparse(got) and got.dtype == np.float32
        assert (got.indices == x.indices).all()
        np.testing.assert_allclose(got.data, ref.data, rtol=2e-7)
    # dense float32 input: in-place float32 semantics
    xd = x[:500].toarray()
    got_d = sb.pp.scale(xd, zero_center=zero_center, max_value=max_value, copy=True)
    ref_d, _, _ = opre.scale(xd, zero_center=zero_center, max_value=max_value)
    assert got_d.dtype == np.float32
    np.testing.assert_allclose(got_d, ref_d, rtol=2e-6, atol=2e-6)
    ad = sb.MiniAnnData(x.copy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sb.pp.scale(ad, zero_center=zero_center, max_value=max_value)
    np.testing.assert_allclose(ad.var["mean"], mean, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(ad.var["std"], std, rtol=1e-9)


# ------------------------------------------------------------------------------------------ louvain / modularity
def test_louvain_vs_networkx_and_planted(graph_small):
    import networkx as nx

    ad, lab = graph_small
    sb.tl.louvain(ad, use_weights=True)
    got = ad.obs["louvain"].to_numpy().astype(int)
    assert ad.uns["louvain"]["params"] == dict(resolution=None, random_state=0)
    conn = ad.obsp["connectivities"]
    q_got = sb.metrics.modularity(conn, got, is_directed=False)
    g = nx.from_scipy_sparse_array(conn)
    parts = nx.community.louvain_communities(g, weight="weight", resolution=1.0, seed=0)
    q_nx = nx.community.modularity(g, parts, weight="weight")
    ref = np.empty(conn.shape[0], int)
    for i, p in enumerate(parts):
        ref[list(p)] = i
    # quality floor: an independent Louvain implementation, and our own modularity kernel against networkx's
    assert abs(sb.metrics.modularity(conn, ref, is_directed=False) - q_nx) < 1e-6
    assert q_got > q_nx - 5e-3, (q_got, q_nx)
    assert adjusted_rand_score(lab, got) > 0.95
    # labels ordered by decreasing size like every scanpy clustering
    sizes = np.bincount(got)
    assert (np.diff(sizes) <= 0).all()
    # same seed -> same labels; restrict_to only relabels the chosen cluster (tests/test_clustering.py:177-213 pattern)
    ad2 = ad.copy()
    sb.tl.louvain(ad2, use_weights=True)
    assert (ad2.obs["louvain"].to_numpy() == ad.obs["louvain"].to_numpy()).all()
    sb.tl.louvain(ad, use_weights=True, restrict_to=("louvain", ["0"]), resolution=2.0)
    r = ad.obs["louvain_R"].astype(str).to_numpy()
    base = ad.obs["louvain"].astype(str).to_numpy()
    assert (r[base != "0"] == base[base != "0"]).all()
    assert all(v.startswith("0,") for v in r[base == "0"])
    with pytest.raises(ValueError, match="flavor"):
        sb.tl.louvain(ad, flavor="nope")


def test_louvain_unweighted_default_and_igraph_flavor(graph_small):
    ad, lab = graph_small
    ad = ad.copy()
    sb.tl.louvain(ad)  # use_weights=False default: every arc weighs 1
    a = ad.obs["louvain"].to_numpy().astype(int)
    assert adjusted_rand_score(lab, a) > 0.9
    with pytest.warns(UserWarning, match="no effect"):
        sb.tl.louvain(ad, flavor="igraph", resolution=3.0, key_added="lv_ig")
    b = ad.obs["lv_ig"].to_numpy().astype(int)
    assert adjusted_rand_score(a, b) > 0.9


def test_metrics_modularity_api(graph_small):
    # tests/test_metrics.py:311-344: in [0, 1] on a clustering, retrieve == calculate == update
    ad, _ = graph_small
    ad = ad.copy()
    sb.tl.leiden(ad, flavor="igraph")
    m_ret = sb.metrics.modularity(ad, mode="retrieve")
    m_calc = sb.metrics.modularity(ad, mode="calculate")
    assert 0 <= m_calc <= 1
    assert abs(m_ret - m_calc) < 1e-9
    ad.uns["leiden"]["modularity"] = -1.0
    assert sb.metrics.modularity(ad, mode="update") == pytest.approx(m_calc)
    assert ad.uns["leiden"]["modularity"] == pytest.approx(m_calc)
    conn = ad.obsp["connectivities"]
    labels = ad.obs["leiden"]
    m_arr = sb.metrics.modularity(conn, labels, is_directed=False)
    assert m_arr == pytest.approx(m_calc)
    assert m_arr == pytest.approx(old.modularity(conn, labels.cat.codes.to_numpy()), abs=1e-7)
    assert sb.metrics.modularity(conn.toarray(), pd.Series(labels.astype(str)), is_directed=False) == pytest.approx(m_calc)
    with pytest.raises(TypeError, match="labels"):
        sb.metrics.modularity(conn, "leiden", is_directed=False)
    with pytest.raises(TypeError, match="is_directed"):
        sb.metrics.modularity(conn, labels)
    with pytest.raises(ValueError, match="undirected"):
        sb.metrics.modularity(ad, is_directed=True)
