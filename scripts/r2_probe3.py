"""Round-2 probe 3: cycles per visit + SM clock under load for the v2 sweep (full / NOSCAN / split-only), clocks via nvidia-smi too."""
import os, sys, time, subprocess, threading
sys.path.insert(0, ".")
import numpy as np, torch
from scanpy_b200 import _ops, _abi
from scanpy_b200._synth import synth_scipy
ctx = _abi.default_context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_300_000
X, lab = synth_scipy(n, 2000)
ip, ix, dat = _ops.csr_to_device(X)
xp = _ops.pca_csr_device(ctx, ip, ix, dat, n, 2000, 50, solver=1)["X_pca"].contiguous()
torch.cuda.synchronize()
os.environ["SB2_KNN2_STAMP"] = "1"
def run(tag, reps=3, **env):
    for k, v in env.items(): os.environ[k] = v
    for r in range(reps):
        idx, dist, info = _ops.knn_device(ctx, xp, 15)
        torch.cuda.synchronize()
    sm = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    print(tag, {k: info[k] for k in ("pass1_ms", "pass1_issued_flops", "n_resweep")}, "issued TF/s", info["pass1_issued_flops"] / info["pass1_ms"] / 1e9, "| after:", sm, flush=True)
    for k in env: os.environ.pop(k)
run("v2 phases", reps=1, SB2_KNN2_NOSCAN="2")
