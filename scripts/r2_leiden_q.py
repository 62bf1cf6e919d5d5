"""Leiden quality on the pbmc68k fixture graph across seeds: CUDA vs the sequential oracle."""
import sys, os
sys.path.insert(0, ".")
import numpy as np
from scipy import sparse
from sklearn.metrics import adjusted_rand_score as ari
from oracle import leiden as old
from scanpy_b200 import _ops
f = np.load("tests/golden/pbmc68k_reduced_graph.npz")
g = sparse.csr_matrix((f["conn_data"].astype(np.float32), f["conn_indices"], f["conn_indptr"]), shape=(700, 700))
orc = [old.leiden(g, seed=s, beta=b) for b in (0.0, 0.01) for s in range(6)]
print("oracle Q", np.round([r[1] for r in orc], 4))
for env in ({}, {"SB2_LEIDEN_EXACT": "1"}):
    os.environ.update(env)
    res = [_ops.leiden(g, seed=s) for s in range(8)]
    print(env, "cuda Q", np.round([r[1] for r in res], 4), "ncomm", [int(r[0].max()) + 1 for r in res],
          "min ARI vs oracle", np.round([min(ari(o[0], r[0]) for o in orc) for r in res], 3), "info", res[0][2])
