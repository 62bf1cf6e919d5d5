#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/r2_probe3.py > gpurun_out/r2_probe3.out 2> gpurun_out/r2_probe3.err
cat gpurun_out/r2_probe3.out; grep "stamp\|phases" gpurun_out/r2_probe3.err | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_sweep2 -s 1 -c 1 -o gpurun_out/r2_knn2_a -f python scripts/r2_knn_once.py > gpurun_out/r2_knn2_a.log 2>&1
tail -3 gpurun_out/r2_knn2_a.log
ls -la gpurun_out/*.ncu-rep
