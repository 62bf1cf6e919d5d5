#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2c_pytest.log 2>&1
tail -8 gpurun_out/r2c_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',{k:d['e2e'][k] for k in ('value','s_per_step')}); print('roofline',{k:d['roofline'].get(k) for k in ('achieved','frac','launch_ms','frac_issued','share_of_step')}); print('stages',{k:v for k,v in d['stages'].items() if k!='parity'}); print('cpu',d.get('cpu_baseline',{}).get('value')); print(d.get('clocks'))
P
timeout 600 python scripts/r2_widened_perf.py > gpurun_out/r2c_widened.log 2>&1; cat gpurun_out/r2c_widened.log | tail -30
timeout 600 python scripts/r2_config_e_share.py > gpurun_out/r2c_config_e.log 2>&1; tail -6 gpurun_out/r2c_config_e.log
MINE='regex:csr_|spmm_|fuzzy_|sym_|knn_prep|knn_tc2_prep|knn_rescore|decide_kernel|agg_insert|lm_apply|dense_sym'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$MINE" -c 400 --csv --log-file gpurun_out/hbm_kernels_r2c.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/bench_under_ncu_r2c.log 2>&1
ls -la gpurun_out | tail -6
