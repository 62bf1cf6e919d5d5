#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
./scripts/ubench/mma_dispatch_stall > gpurun_out/r2_ubench_dispatch.txt 2>&1
cat gpurun_out/r2_ubench_dispatch.txt
timeout 600 python scripts/r2_probe1.py > gpurun_out/r2_probe1.out 2> gpurun_out/r2_probe1.err
tail -5 gpurun_out/r2_probe1.out
