"""sb2_csr_gram at the bench workload: one-RED-per-product kernel (default) vs the tiled shared-memory variant (SB2_GRAM_TILED=1 in a second process), result
equality and CUDA-event timings.  usage: python scripts/r2_gram.py [n]"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from scanpy_b200 import _abi, _ops
from scanpy_b200._abi import check, ptr
from scanpy_b200._synth import synth_scipy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
g = 2000
X, _ = synth_scipy(n, g)
ctx = _abi.default_context()
ip, ix, dat = _ops.csr_to_device(X)
G = torch.empty((g, g), dtype=torch.float64, device="cuda")
def run():
    check(ctx.lib.sb2_csr_gram(ctx.handle, n, g, ptr(ip), ptr(ix), ptr(dat), ptr(G)))
for _ in range(2): run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
tag = "tiled" if os.environ.get("SB2_GRAM_TILED") else "v1"
print(f"gram {tag}: n={n} min {min(ts):.2f} ms median {sorted(ts)[2]:.2f} ms; checksum {float(G.sum()):.10e} trace {float(G.diagonal().sum()):.10e}")
if n <= 200_000:
    ref = (X.astype(np.float64).T @ X.astype(np.float64)).toarray()
    print("max rel err vs scipy fp64:", float(np.abs(G.cpu().numpy() - ref).max() / np.abs(ref).max()))
