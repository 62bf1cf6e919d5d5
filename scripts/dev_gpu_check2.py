"""Throw-away GPU check #2: fuzzy set + Leiden vs oracle; big timings."""
import json, sys, time, os
import numpy as np
sys.path.insert(0, ".")
import torch
from scanpy_b200 import _abi, _ops
from scanpy_b200._synth import synth_scipy
from oracle import knn as oknn, pca as opca, fuzzy as ofz, leiden as old
from sklearn.metrics import adjusted_rand_score as ari
from scipy import sparse
out = {}
ctx = _abi.default_context()
L = np.load("tests/golden/reference_test_literals.npz")
f = np.load("tests/golden/pbmc68k_reduced_graph.npz")
# fuzzy goldens
idx, dist, _ = _ops.knn(L["X4"].astype(np.float32), 3)
c, s, r = _ops.fuzzy_simplicial_set(idx, dist)
print("fuzzy 4pt err", np.abs(c.toarray() - L["connectivities_umap"]).max(), flush=True)
n, k = 700, int(f["n_neighbors"][0])
di = f["dist_indices"].reshape(n, k - 1); dd = f["dist_data"].reshape(n, k - 1)
o = np.argsort(dd, axis=1, kind="stable"); di = np.take_along_axis(di, o, 1); dd = np.take_along_axis(dd, o, 1)
idx = np.hstack([np.arange(n)[:, None], di]).astype(np.int32); dist = np.hstack([np.zeros((n, 1)), dd])
c, s, r = _ops.fuzzy_simplicial_set(idx, dist)
g = sparse.csr_matrix((f["conn_data"], f["conn_indices"], f["conn_indptr"]), shape=(n, n)); g.sort_indices()
print("fuzzy fixture nnz", c.nnz, g.nnz, "pattern", bool((c.indices == g.indices).all() and (c.indptr == g.indptr).all()),
      "maxerr", float(np.abs(c.data - g.data).max()), "sorted", bool(c.has_sorted_indices), flush=True)
oc, os_, or_ = ofz.fuzzy_simplicial_set(idx, dist, n, k)
print("vs oracle", float(abs(c - oc).max()), float(np.abs(s - os_).max()), float(np.abs(r - or_).max()), flush=True)
# leiden on fixture graph
m, q, info = _ops.leiden(g.astype(np.float32), seed=0)
mo, qo, _ = old.leiden(g, seed=0)
print("leiden fixture: gpu Q", q, "ncomm", m.max() + 1, info, "oracle Q", qo, mo.max() + 1, "ARI", ari(m, mo),
      "Q check", old.modularity(g, m), _ops.modularity(g.astype(np.float32), m), flush=True)
m2, q2, _ = _ops.leiden(g.astype(np.float32), seed=0)
print("determinism", bool((m == m2).all()), q == q2, flush=True)
# synthetic pipeline at 20k and 100k
for nn in (20000, 100000):
    X, lab = synth_scipy(nn, 2000, device="cuda")
    t = time.time(); P = _ops.pca_csr(X, 50, solver=1); t_pca = time.time() - t
    t = time.time(); idx, dist, kinfo = _ops.knn(P["X_pca"], 15); t_knn = time.time() - t
    t = time.time(); C, s, r = _ops.fuzzy_simplicial_set(idx, dist); t_fz = time.time() - t
    t = time.time(); m, q, info = _ops.leiden(C, seed=0); t_ld = time.time() - t
    t = time.time(); oC, _, _ = ofz.fuzzy_simplicial_set(idx, dist, nn, 15); t_ofz = time.time() - t
    t = time.time(); mo, qo, po = old.leiden(oC, seed=0); t_old = time.time() - t
    res = dict(n=nn, t_pca=t_pca, t_knn=t_knn, t_fuzzy=t_fz, t_leiden=t_ld, t_oracle_fuzzy=t_ofz, t_oracle_leiden=t_old,
               fuzzy_maxerr=float(abs(C - oC).max()), fuzzy_nnz=(C.nnz, oC.nnz), Q=q, Qo=qo, ncomm=int(m.max() + 1), ncomm_o=int(mo.max() + 1),
               ari_vs_oracle=ari(m, mo), ari_planted=ari(lab, m), ari_oracle_planted=ari(lab, mo), info=info, kinfo=kinfo)
    print(res, flush=True)
    out[f"pipe_{nn}"] = res
# big timings: 1.3M
nn = 1_300_000
t = time.time(); X, lab = synth_scipy(nn, 2000, device="cuda"); print("gen 1.3M", time.time() - t, X.nnz, flush=True)
d = _ops.csr_to_device(X); torch.cuda.synchronize()
for solver in (1,):
    torch.cuda.synchronize(); t = time.time()
    P = _ops.pca_csr_device(ctx, *d, nn, 2000, 50, solver=solver)
    torch.cuda.synchronize(); dt = time.time() - t
    print("pca 1.3M solver", solver, dt, P["iterations"], P["converged"], P["max_rel_residual"], flush=True)
    out[f"pca_1.3M_s{solver}"] = dict(sec=dt, it=P["iterations"])
xp = P["X_pca"]
torch.cuda.synchronize(); t = time.time()
idx, dist, kinfo = _ops.knn_device(ctx, xp, 15)
torch.cuda.synchronize(); dt = time.time() - t
print("knn 1.3M", dt, 2 * nn * nn * 50 / dt / 1e12, kinfo, flush=True)
out["knn_1.3M"] = dict(sec=dt, tflops=2 * nn * nn * 50 / dt / 1e12, info=kinfo)
torch.cuda.synchronize(); t = time.time()
ip, ii, dd_, sg, rh = _ops.fuzzy_simplicial_set_device(ctx, idx, dist, nn, 15)
torch.cuda.synchronize(); dt = time.time() - t
print("fuzzy 1.3M", dt, ii.numel(), flush=True); out["fuzzy_1.3M"] = dict(sec=dt, nnz=int(ii.numel()))
torch.cuda.synchronize(); t = time.time()
mem, q, nc, info = _ops.leiden_device(ctx, ip, ii, dd_, nn, seed=0)
torch.cuda.synchronize(); dt = time.time() - t
mm = mem.cpu().numpy()
print("leiden 1.3M", dt, q, nc, info, "ARI planted", ari(lab, mm), flush=True)
out["leiden_1.3M"] = dict(sec=dt, Q=q, nc=nc, info=info, ari_planted=ari(lab, mm))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dev_check2.json", "w"), indent=1, default=str)
print("DONE")
