"""Throw-away GPU check #4: Chebyshev PCA parity/speed; stage timings of the e2e API; Leiden timing."""
import sys, time, os
import numpy as np
sys.path.insert(0, ".")
import torch
import scanpy_b200 as sb
from scanpy_b200 import _abi, _ops
from scanpy_b200._synth import synth_scipy
from oracle import pca as opca
ctx = _abi.default_context()
def relerr(a, b):
    a = opca.align_signs(np.asarray(a, np.float64), b)
    return np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0)
for (n, g, k) in [(3000, 500, 20), (20000, 2000, 50)]:
    X, lab = synth_scipy(n, g, device="cuda")
    o64 = opca.pca_arpack(X.astype(np.float64), k, dtype="float64")
    for solver in (1, 0):
        torch.cuda.synchronize(); t = time.time()
        r = _ops.pca_csr(X, k, solver=solver)
        torch.cuda.synchronize(); dt = time.time() - t
        e = relerr(r["X_pca"], o64["X_pca"])
        print(dict(n=n, g=g, solver=solver, sec=round(dt, 4), it=r["iterations"], conv=r["converged"], res=r["max_rel_residual"], err_max=float(e.max()), err_med=float(np.median(e))), flush=True)
nn = 1_300_000
X, lab = synth_scipy(nn, 2000, device="cuda")
d = _ops.csr_to_device(X); torch.cuda.synchronize()
for solver in (1, 0):
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        P = _ops.pca_csr_device(ctx, *d, nn, 2000, 50, solver=solver)
        torch.cuda.synchronize(); dt = time.time() - t
    print("pca 1.3M solver", solver, round(dt, 4), P["iterations"], P["converged"], P["max_rel_residual"], flush=True)
    if solver == 1: P1 = P
e = relerr(P["X_pca"][:200000].cpu().numpy(), P1["X_pca"][:200000].cpu().numpy().astype(np.float64))
print("spmm vs gram X_pca rel err (first 200k rows): max", float(e.max()), "median", float(np.median(e)), flush=True)
# e2e stage timings
os.environ["SB2_TIMING"] = "1"
ad = sb.MiniAnnData(X)
for rep in range(2):
    ad = sb.MiniAnnData(X)
    t0 = time.time(); sb.pp.pca(ad, n_comps=50); torch.cuda.synchronize(); t1 = time.time()
    sb.pp.neighbors(ad, n_neighbors=15); torch.cuda.synchronize(); t2 = time.time()
    sb.tl.leiden(ad, flavor="igraph", n_iterations=-1); torch.cuda.synchronize(); t3 = time.time()
    print("e2e stages: pca", round(t1 - t0, 3), "neighbors", round(t2 - t1, 3), "leiden", round(t3 - t2, 3), flush=True)
import cProfile, pstats
ad = sb.MiniAnnData(X)
pr = cProfile.Profile(); pr.enable()
sb.pp.pca(ad, n_comps=50); sb.pp.neighbors(ad, n_neighbors=15); sb.tl.leiden(ad, flavor="igraph", n_iterations=-1)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
print("DONE")
