#!/bin/bash
# 2-GPU sanity of the sharded path (gpurun --gpus 2): the bench line at N = 2 with its parity block (labels_sha1 must equal the N = 1 run's)
mkdir -p gpurun_out
SB2_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e',{k:d['e2e'][k] for k in ('value','s_per_step')}); print('roofline',{k:d['roofline'][k] for k in ('frac','launch_ms','share_of_step')}); print('parity',d['stages'].get('parity'))
P
grep "sb2 leiden\] n=" gpurun_out/r2_bench_n2.err | tail -2; tail -3 gpurun_out/r2_bench_n2.err | cut -c1-300
