"""Round-2 probe: kNN sweep cycle stamps (SB2_KNN_DBG modes) + per-phase timings at the bench workload."""
import os, sys, time
sys.path.insert(0, ".")
import torch
from scanpy_b200 import _ops, _abi
from scanpy_b200._synth import synth_scipy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
X, lab = synth_scipy(n, 2000)
ctx = _abi.default_context()
ip, ix, dat = _ops.csr_to_device(X)
xp = _ops.pca_csr_device(ctx, ip, ix, dat, n, 2000, 50, solver=1)["X_pca"].contiguous()
torch.cuda.synchronize()
for rep in range(2):
    idx, dist, info = _ops.knn_device(ctx, xp, 15)
    torch.cuda.synchronize()
    print("plain", info, flush=True)
for mode in ("1", "2", "3"):
    os.environ["SB2_KNN_DBG"] = mode
    sys.stderr.write(f"==== SB2_KNN_DBG={mode}\n"); sys.stderr.flush()
    idx, dist, info = _ops.knn_device(ctx, xp, 15)
    torch.cuda.synchronize()
    print("dbg", mode, info, flush=True)
os.environ.pop("SB2_KNN_DBG")
os.environ["SB2_TIMING"] = "1"
out = _ops.pca_csr_device(ctx, ip, ix, dat, n, 2000, 50, solver=1)
idx, dist, info = _ops.knn_device(ctx, xp, 15)
c = _ops.fuzzy_simplicial_set_device(ctx, idx, dist, n, 15)
t = time.perf_counter()
m = _ops.leiden_device(ctx, c[0], c[1], c[2], n)
torch.cuda.synchronize()
print("leiden wall (SB2_TIMING on)", time.perf_counter() - t, m[1:], flush=True)
