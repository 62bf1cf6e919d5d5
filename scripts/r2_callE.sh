#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_widened.py -m gpu -q -k "umap" 2>&1 | tail -3
timeout 600 python scripts/r2_widened_perf.py 2>&1 | tee gpurun_out/r2e_widened.log | grep -E "umap|embedding|silhouette"
