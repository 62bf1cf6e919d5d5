#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
grep "bench " gpurun_out/r2_bench.err | tail -8
cat gpurun_out/r2_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',{k:d['e2e'][k] for k in ('value','s_per_step','h2d_bytes_per_step','d2h_bytes_per_step')}); print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','launch_ms','frac_issued','share_of_step')}); print('stages',d['stages']); print('cpu',d.get('cpu_baseline')); print(d.get('clocks'))"
tail -3 gpurun_out/r2_bench.err
