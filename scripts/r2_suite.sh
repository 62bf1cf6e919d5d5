#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2_pytest.log
tail -25 gpurun_out/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cat gpurun_out/r2_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',d['e2e']); print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','launch_ms','frac_issued','share_of_step')}); print('stages',d['stages']); print('cpu',d.get('cpu_baseline',{}).get('value'), d.get('clocks'))"
tail -5 gpurun_out/r2_bench.err
