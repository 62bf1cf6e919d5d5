"""Per-stage device times of one pipeline step at the bench workload (CUDA events), plus Leiden's phase log."""
import os, sys, time
sys.path.insert(0, ".")
import torch
from scanpy_b200 import _ops, _abi
from scanpy_b200._synth import synth_scipy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
X, lab = synth_scipy(n, 2000)
ctx = _abi.default_context()
ip, ix, dat = _ops.csr_to_device(X)
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
for rep in range(3):
    if rep == 2: os.environ["SB2_TIMING"] = "1"
    e0 = ev(); p = _ops.pca_csr_device(ctx, ip, ix, dat, n, 2000, 50, solver=1)
    e1 = ev(); idx, dist, info = _ops.knn_device(ctx, p["X_pca"], 15)
    e2 = ev(); c = _ops.fuzzy_simplicial_set_device(ctx, idx, dist, n, 15)
    e3 = ev(); m = _ops.leiden_device(ctx, c[0], c[1], c[2], n)
    e4 = ev(); torch.cuda.synchronize()
    print(f"rep {rep}: pca {e0.elapsed_time(e1):.1f} ms | knn {e1.elapsed_time(e2):.1f} (sweep {info['pass1_ms']:.1f}) | fuzzy {e2.elapsed_time(e3):.1f} | leiden {e3.elapsed_time(e4):.1f} | total {e0.elapsed_time(e4):.1f}; Q {m[1]:.5f} ncomm {m[2]} {m[3]}", flush=True)
