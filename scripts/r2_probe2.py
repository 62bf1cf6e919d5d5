"""Round-2 probe 2: gen-2 tensor sweep: exact vs fp64 oracle on assorted shapes; gen 1 == gen 2 at the bench workload; timings; NOSCAN floor."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from scanpy_b200 import _ops, _abi
from scanpy_b200._synth import synth_scipy
ctx = _abi.default_context()
from oracle import knn as oknn
for (n, d, k) in [(5, 2, 5), (1000, 50, 15), (4097, 33, 10), (12345, 50, 30), (3000, 150, 56), (2500, 74, 15), (3000, 100, 30), (20000, 120, 15), (300_000, 12, 15)]:
    rs = np.random.RandomState(n + d)
    x = rs.standard_normal((n, d)).astype(np.float32); x[: n // 3] += 2.5
    t = time.perf_counter(); idx, dist, info = _ops.knn(x, k); dt = time.perf_counter() - t
    rows = np.arange(n) if n <= 20000 else rs.choice(n, 2000, replace=False)
    oi, od = oknn.knn_exact_f64(x, rows, k)
    bad = oknn.exact_set_mismatches(idx[rows], oi, od, k).sum()
    print(f"n={n} d={d} k={k}: mismatching rows {bad}/{len(rows)} gen={info['pass1_tensor']} resweep={info['n_resweep']} uncert={info['n_uncertified']} pass1 {info['pass1_ms']:.3f} ms, call {dt:.3f}s", flush=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
X, lab = synth_scipy(n, 2000)
ip, ix, dat = _ops.csr_to_device(X)
xp = _ops.pca_csr_device(ctx, ip, ix, dat, n, 2000, 50, solver=1)["X_pca"].contiguous()
torch.cuda.synchronize()
os.environ["SB2_KNN2_STAMP"] = "1"
res = {}
for gen in ("2", "1"):
    os.environ["SB2_KNN_V"] = gen
    for rep in range(2):
        idx, dist, info = _ops.knn_device(ctx, xp, 15)
        torch.cuda.synchronize()
    print("gen", gen, info, "issued TF/s", info["pass1_issued_flops"] / info["pass1_ms"] / 1e9, flush=True)
    res[gen] = (idx.clone(), dist.clone())
print("v1 == v2 idx:", bool((res["1"][0] == res["2"][0]).all()), "dist:", bool((res["1"][1] == res["2"][1]).all()), flush=True)
os.environ["SB2_KNN_V"] = "2"
def run(tag, k=15, **env):
    for kk, v in env.items(): os.environ[kk] = v
    idx, dist, info = _ops.knn_device(ctx, xp, k); torch.cuda.synchronize()
    print(tag, info, "issued TF/s", info["pass1_issued_flops"] / info["pass1_ms"] / 1e9, "same idx:", bool((idx == res["2"][0]).all()) if k == 15 else None, flush=True)
    for kk in env: os.environ.pop(kk)
run("gen 2 NOSCAN", SB2_KNN2_NOSCAN="1")
run("gen 2 cold start", SB2_KNN_EST="0")
run("gen 2 split-only", SB2_KNN_TIERS="3")
run("gen 2 k=30", k=30)
