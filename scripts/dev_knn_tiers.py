"""Throw-away: tiered kNN on the bench workload's own X_pca (1.3M x 50): sweep time and re-sweep counts per setting."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from scanpy_b200 import _ops, _abi
from scanpy_b200._synth import synth_scipy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
X, lab = synth_scipy(n, 2000)
ctx = _abi.default_context()
ip, ix, dat = _ops.csr_to_device(X)
res = _ops.pca_csr_device(ctx, ip, ix, dat, n, 2000, 50, solver=1)
xp = res["X_pca"].contiguous()
print("x_pca", xp.shape, "max norm", float(xp.norm(dim=1).max()), "mean norm", float(xp.norm(dim=1).mean()), flush=True)
ref = None
for env in [dict(SB2_KNN_TIERS="3"), dict(), dict(SB2_KNN_LIST="64"), dict(SB2_KNN_TIERS="3", SB2_KNN_LIST="64")]:
    for k_ in ("SB2_KNN_TIERS", "SB2_KNN_LIST"):
        os.environ.pop(k_, None)
    os.environ.update(env)
    for rep in range(2):
        idx, dist, info = _ops.knn_device(ctx, xp, 15)
    torch.cuda.synchronize()
    same = None if ref is None else bool((idx == ref).all())
    ref = idx if ref is None else ref
    print(env, {kk: info[kk] for kk in ("pass1_ms", "n_uncertified", "n_resweep")}, "issued TF/s %.0f" % (info["pass1_issued_flops"] / info["pass1_ms"] / 1e9), "same as first:", same, flush=True)
