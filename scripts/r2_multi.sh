#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
grep "bench " gpurun_out/r2_bench_n$N.err | tail -4
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e',{k:d['e2e'][k] for k in ('value','s_per_step')}); print('roofline',{k:d['roofline'][k] for k in ('frac','launch_ms','share_of_step')}); print('parity',d['stages'].get('parity'))"
tail -3 gpurun_out/r2_bench_n$N.err | cut -c1-300
