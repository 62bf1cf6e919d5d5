"""Throw-away GPU check #5: Leiden phase statistics at 1.3M."""
import sys, time, os
import numpy as np
sys.path.insert(0, ".")
import torch
from scanpy_b200 import _abi, _ops
from scanpy_b200._synth import synth_scipy
ctx = _abi.default_context()
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
X, lab = synth_scipy(nn, 2000, device="cuda")
d = _ops.csr_to_device(X)
P = _ops.pca_csr_device(ctx, *d, nn, 2000, 50, solver=1)
idx, dist, kinfo = _ops.knn_device(ctx, P["X_pca"], 15)
ip, ii, dd_, sg, rh = _ops.fuzzy_simplicial_set_device(ctx, idx, dist, nn, 15)
torch.cuda.synchronize()
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    mem, q, nc, info = _ops.leiden_device(ctx, ip, ii, dd_, nn, seed=0)
    torch.cuda.synchronize(); print("leiden", time.time() - t, q, nc, info, flush=True)
os.environ["SB2_TIMING"] = "1"
mem, q, nc, info = _ops.leiden_device(ctx, ip, ii, dd_, nn, seed=0)
from sklearn.metrics import adjusted_rand_score
print("ARI planted", adjusted_rand_score(lab, mem.cpu().numpy()))
