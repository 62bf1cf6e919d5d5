"""Throw-away: tensor-core kNN at wide embeddings (d = 50 / 64 / 100 / 150), timing + certification rate."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from scanpy_b200 import _ops, _abi
rs = np.random.RandomState(0)
n = 325_000
for d, k in [(50, 15), (57, 15), (64, 15), (73, 15), (100, 15), (100, 30), (150, 15)]:
    c = rs.standard_normal((24, d)).astype(np.float32) * 3
    x = torch.from_numpy((c[rs.randint(0, 24, n)] + rs.standard_normal((n, d)) * (0.97 ** np.arange(d))).astype(np.float32)).cuda()
    for rep in range(2):
        idx, dist, info = _ops.knn_device(_abi.default_context(), x, k)
    print(d, k, {kk: info[kk] for kk in ("pass1_ms", "n_uncertified", "n_resweep")},
          "alg TF/s %.0f issued TF/s %.0f" % (info["pass1_flops"] / info["pass1_ms"] / 1e9, info["pass1_issued_flops"] / info["pass1_ms"] / 1e9), flush=True)
