"""Config E (10M cells x 4000 genes, n_pcs 100, k 30, 8 GPUs), ONE RANK'S SHARE on one B200: the kNN of 1.25M query rows
against all 10M points (d = 100, k = 30), with memory accounting - the step that dominates config E and the one whose
resident operand set (all points + their tensor-core images) decides whether E fits 180 GB per GPU.
usage: python scripts/r2_config_e_share.py [n_points] [n_query]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from scanpy_b200 import _abi, _ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1_250_048
d, k = 100, 30
torch.manual_seed(0)
free0, total = torch.cuda.mem_get_info()
cent = torch.randn(64, d, device="cuda") * 3.0
X = torch.empty((n, d), dtype=torch.float32, device="cuda")
for s in range(0, n, 1_000_000):
    e = min(n, s + 1_000_000)
    lab = torch.randint(0, 64, (e - s,), device="cuda")
    X[s:e] = cent[lab] + torch.randn(e - s, d, device="cuda")
torch.cuda.synchronize()
ctx = _abi.default_context()
free1, _ = torch.cuda.mem_get_info()
print(f"points {n} x {d}: X {X.numel()*4/1e9:.2f} GB; device {total/1e9:.0f} GB, free before kNN {free1/1e9:.1f} GB", flush=True)
lo = [free1]
t0 = time.time()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
idx, dist, info = _ops.knn_device(ctx, X, k, q0=0, n_query=nq)
e1.record(); torch.cuda.synchronize()
free2, _ = torch.cuda.mem_get_info()
ms = e0.elapsed_time(e1)
fl = 2.0 * nq * n * d
print(f"kNN {nq} queries x {n} points, d={d}, k={k}: {ms:.0f} ms total, sweep {info['pass1_ms']:.0f} ms = {fl/info['pass1_ms']/1e9:.0f} TFLOP/s algorithmic "
      f"({info['pass1_issued_flops']/info['pass1_ms']/1e9:.0f} issued); resweep rows {info['n_resweep']}, uncertified {info['n_uncertified']}, tensor path {info['pass1_tensor']}")
print(f"memory: pool high-water (free before - free after, scratch stays pooled) {(free1-free2)/1e9:.1f} GB; per-rank resident for E: X_pca {n*d*4/1e9:.1f} GB + "
      f"outputs {nq*k*12/1e9:.2f} GB")
# sampled exactness: 256 query rows brute-forced in fp64 on the device
rows = torch.randint(0, nq, (256,), device="cuda")
q = X[rows].double()
best = torch.full((256, k), float("inf"), dtype=torch.float64, device="cuda"); bi = torch.zeros((256, k), dtype=torch.int64, device="cuda")
for s in range(0, n, 2_000_000):
    e = min(n, s + 2_000_000)
    dd = torch.cdist(q, X[s:e].double()) ** 2
    allv = torch.cat([best, dd], 1); alli = torch.cat([bi, torch.arange(s, e, device="cuda").expand(256, -1)], 1)
    v, o = torch.topk(allv, k, dim=1, largest=False); best, bi = v, torch.gather(alli, 1, o)
got = idx[rows].long()
same = sum(set(a.tolist()) == set(b.tolist()) for a, b in zip(got, bi))
print(f"sampled rows with identical neighbour sets vs fp64 brute force: {same}/256")
