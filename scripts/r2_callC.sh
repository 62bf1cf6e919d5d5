#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_widened.py -m gpu -q -k "overlapped or chunked or zarr" 2>&1 | tail -5
for ov in 0 1; do
  SB2_PCA_OVERLAP=$ov timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2d_bench_ov$ov.json 2> gpurun_out/r2d_bench_ov$ov.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r2d_bench_ov$ov.json').read().strip().splitlines()[-1])
print('overlap=$ov', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', {k:d['e2e'][k] for k in ('value','s_per_step','h2d_bytes_per_step','d2h_bytes_per_step')})
P
done
# where the e2e seconds go: wall time of each public call of one e2e step (pinned inputs), overlap on
python - <<'P'
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import scanpy_b200 as sb
from scipy import sparse
from scanpy_b200._synth import synth_scipy
x, _ = synth_scipy(1_300_000, 2000)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
xp = sparse.csr_matrix((pin(x.data), pin(x.indices), pin(x.indptr)), shape=x.shape, copy=False)
ad = sb.MiniAnnData(xp)
for rep in range(3):
    ts = []
    for f in (lambda: sb.pp.pca(ad, n_comps=50), lambda: sb.pp.neighbors(ad, n_neighbors=15), lambda: sb.tl.leiden(ad, flavor="igraph", n_iterations=-1)):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    print("e2e wall per call: pca %.3f neighbors %.3f leiden %.3f total %.3f" % (*ts, sum(ts)), flush=True)
P
