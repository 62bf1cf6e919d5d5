"""One kNN call at the bench workload (for ncu)."""
import sys
sys.path.insert(0, ".")
import torch
from scanpy_b200 import _ops, _abi
from scanpy_b200._synth import synth_scipy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
X, lab = synth_scipy(n, 2000)
ctx = _abi.default_context()
ip, ix, dat = _ops.csr_to_device(X)
xp = _ops.pca_csr_device(ctx, ip, ix, dat, n, 2000, 50, solver=1)["X_pca"].contiguous()
idx, dist, info = _ops.knn_device(ctx, xp, 15)
torch.cuda.synchronize()
print(info)
