#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:knn_pass1_tc -c 1 -o gpurun_out/prof_knn_tc \
    python - <<'PY' > gpurun_out/prof_tc.log 2>&1
import sys; sys.path.insert(0, ".")
import torch
from scanpy_b200 import _abi, _ops
ctx = _abi.default_context()
x = torch.randn(325000, 50, device="cuda") * 3
idx, dist, info = _ops.knn_device(ctx, x, 15)
print(info)
PY
ls -la gpurun_out
