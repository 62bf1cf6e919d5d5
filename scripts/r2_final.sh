#!/bin/bash
# final round-2 verification as the driver runs it: gpu tests, smoke, bench (driver flags)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r2f_pytest.log 2>&1
tail -8 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',{k:d['e2e'][k] for k in ('value','s_per_step')}); print('roofline',{k:d['roofline'].get(k) for k in ('achieved','frac','launch_ms','frac_issued','share_of_step','traffic')}); print('stages',{k:v for k,v in d['stages'].items() if k!='parity'}); print('parity',d['stages'].get('parity')); print('cpu',d.get('cpu_baseline',{}).get('value')); print(d.get('clocks'))
P
tail -2 gpurun_out/r2f_bench.err | cut -c1-300
