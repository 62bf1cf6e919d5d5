"""Throw-away GPU check (first contact): kNN + PCA vs the oracle on small inputs, timings at 100k."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from scanpy_b200 import _abi, _ops
from scanpy_b200._synth import synth_scipy
from oracle import knn as oknn, pca as opca

out = {}
ctx = _abi.default_context()
di = ctx.device_info()
print("device", di.name, di.sm_count, di.clock_khz)
rs = np.random.RandomState(0)

def t_knn(n, d, k, check=True):
    x = rs.standard_normal((n, d)).astype(np.float32)
    x[: n // 2] += 3.0
    torch.cuda.synchronize(); t = time.time()
    idx, dist, info = _ops.knn(x, k)
    torch.cuda.synchronize(); dt = time.time() - t
    res = dict(n=n, d=d, k=k, sec=dt, info=info)
    if check:
        oi, od = oknn.knn_brute(x, k)
        same = oknn.same_neighbor_sets(idx, dist, oi, od)
        res["rows_equal"] = int(same.sum()); res["max_dist_err"] = float(np.abs(dist - od).max())
        res["self_col0"] = bool((idx[:, 0] == np.arange(n)).all())
    print(res, flush=True)
    return res

out["knn_small"] = [t_knn(n, d, k) for (n, d, k) in [(4, 2, 3), (100, 5, 10), (1000, 50, 15), (5000, 50, 15), (20000, 30, 15)]]
# duplicates / ties
x = np.zeros((300, 10), np.float32); x[100:] = rs.standard_normal((200, 10))
idx, dist, info = _ops.knn(x, 15)
oi, od = oknn.knn_brute(x, 15)
print("dups: info", info, "dist equal", np.allclose(dist, od, atol=1e-6), flush=True)
out["knn_dups"] = dict(info=info, dist_equal=bool(np.allclose(dist, od, atol=1e-6)))
# device-only timing at 100k and 300k
for n in (100_000, 300_000):
    x = torch.randn(n, 50, device="cuda")
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        idx, dist, info = _ops.knn_device(ctx, x, 15)
        torch.cuda.synchronize(); dt = time.time() - t
    print("knn device", n, dt, "TFLOP/s", 2 * n * n * 50 / dt / 1e12, info, flush=True)
    out[f"knn_dev_{n}"] = dict(sec=dt, tflops=2 * n * n * 50 / dt / 1e12, info=info)

# PCA
def t_pca(n, g, k, solver):
    X, lab = synth_scipy(n, g, device="cuda")
    torch.cuda.synchronize(); t = time.time()
    r = _ops.pca_csr(X, k, solver=solver)
    torch.cuda.synchronize(); dt = time.time() - t
    o = opca.pca_arpack(X, k)
    o64 = opca.pca_arpack(X.astype(np.float64), k, dtype="float64")
    def relerr(a, b):
        a = opca.align_signs(a, b)
        return np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0)
    e32 = relerr(r["X_pca"], o["X_pca"]); e64 = relerr(r["X_pca"], o64["X_pca"]); eref = relerr(o["X_pca"].astype(np.float64), o64["X_pca"])
    s = o64["singular_values"]; gaps = (s[:-1] - s[1:]) / s[:-1]
    res = dict(n=n, g=g, k=k, solver=solver, sec=dt, it=r["iterations"], conv=r["converged"], res=r["max_rel_residual"],
               err_vs_ref32_max=float(e32.max()), err_vs_ref64_max=float(e64.max()), ref32_vs_ref64_max=float(eref.max()),
               err_vs_ref64_med=float(np.median(e64)), min_gap=float(gaps.min()),
               var_rel=float(np.abs(r["variance"] / o64["variance"] - 1).max()),
               ratio_rel=float(np.abs(r["variance_ratio"] / o64["variance_ratio"] - 1).max()),
               sign_ok=bool((np.sign(np.einsum("ij,ij->j", r["X_pca"], o64["X_pca"])) > 0).all()))
    print(res, flush=True)
    return res

out["pca"] = []
for (n, g, k) in [(3000, 500, 20), (20000, 2000, 50)]:
    for solver in (1, 0):
        try:
            out["pca"].append(t_pca(n, g, k, solver))
        except Exception as e:
            print("PCA FAIL", n, g, k, solver, repr(e), flush=True)
            out["pca"].append(dict(n=n, g=g, solver=solver, error=repr(e)))
# golden A_list
L = np.load("tests/golden/reference_test_literals.npz")
from scipy import sparse
try:
    r = _ops.pca_csr(sparse.csr_matrix(L["A_list"].astype(np.float32)), 4, solver=1)
    print("A_pca golden norm", np.linalg.norm(np.abs(L["A_pca"][:, :4]) - np.abs(r["X_pca"])), flush=True)
    out["pca_golden"] = float(np.linalg.norm(np.abs(L["A_pca"][:, :4]) - np.abs(r["X_pca"])))
except Exception as e:
    print("golden FAIL", repr(e)); out["pca_golden"] = repr(e)
# PCA timing at 100k x 2000
X, lab = synth_scipy(100_000, 2000, device="cuda")
d = _ops.csr_to_device(X)
for solver in (1, 0):
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        r = _ops.pca_csr_device(ctx, *d, X.shape[0], X.shape[1], 50, solver=solver)
        torch.cuda.synchronize(); dt = time.time() - t
    print("pca device 100k solver", solver, dt, r["iterations"], r["converged"], r["max_rel_residual"], flush=True)
    out[f"pca_dev_100k_s{solver}"] = dict(sec=dt, it=r["iterations"], conv=r["converged"], res=r["max_rel_residual"])
import os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dev_check1.json", "w"), indent=1, default=str)
print("DONE")
