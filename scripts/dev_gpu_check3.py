"""Throw-away GPU check #3: tensor-core kNN pass vs oracle, timings."""
import json, sys, time, os
import numpy as np
sys.path.insert(0, ".")
import torch
from scanpy_b200 import _abi, _ops
from oracle import knn as oknn
ctx = _abi.default_context()
rs = np.random.RandomState(0)
def t_knn(n, d, k):
    x = rs.standard_normal((n, d)).astype(np.float32) * 3
    x[: n // 2] += 5.0
    idx, dist, info = _ops.knn(x, k)
    oi, od = oknn.knn_brute(x, k)
    same = oknn.same_neighbor_sets(idx, dist, oi, od)
    print(dict(n=n, d=d, k=k, rows_equal=int(same.sum()), info=info, maxerr=float(np.abs(dist[:,1:] - od[:,1:]).max())), flush=True)
for (n, d, k) in [(4, 2, 3), (100, 5, 10), (300, 50, 15), (1000, 50, 15), (5000, 50, 15), (20000, 30, 15), (3000, 52, 30), (2000, 60, 8)]:
    try:
        t_knn(n, d, k)
    except Exception as e:
        print("FAIL", n, d, k, repr(e), flush=True)
x = np.zeros((600, 12), np.float32); x[200:] = rs.standard_normal((400, 12)); x[300:340] = x[299]
idx, dist, info = _ops.knn(x, 15); oi, od = oknn.knn_brute(x, 15)
print("dups", info, np.allclose(dist, od, atol=1e-6), flush=True)
for n in (100_000, 325_000, 1_300_000):
    x = torch.randn(n, 50, device="cuda") * 3
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        idx, dist, info = _ops.knn_device(ctx, x, 15)
        torch.cuda.synchronize(); dt = time.time() - t
    print("knn device", n, dt, "pass1 ms", info["pass1_ms"], "TFLOP/s(2nnd)", info["pass1_flops"] / info["pass1_ms"] / 1e9, info, flush=True)
os.environ["SB2_KNN_PASS1"] = "ffma"
x = torch.randn(325_000, 50, device="cuda") * 3
idx2, dist2, info2 = _ops.knn_device(ctx, x, 15)
os.environ.pop("SB2_KNN_PASS1")
idx1, dist1, info1 = _ops.knn_device(ctx, x, 15)
print("tc vs ffma identical idx:", bool((idx1 == idx2).all()), "dist:", bool((dist1 == dist2).all()), info1, info2, flush=True)
print("DONE")
