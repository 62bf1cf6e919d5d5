#!/bin/bash
# Round-2 GPU pass: tests, smoke, bench, stage times, ncu launch list + full capture of the dominant kernel.
# Numbers printed under ncu are NOT bench values.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r2_pytest.log 2>&1
tail -6 gpurun_out/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',{k:d['e2e'][k] for k in ('value','s_per_step')}); print('roofline',{k:d['roofline'].get(k) for k in ('achieved','frac','launch_ms','frac_issued','share_of_step')}); print('stages',d['stages']); print('cpu',d.get('cpu_baseline',{}).get('value')); print(d.get('clocks'))
P
tail -3 gpurun_out/r2_bench.err | cut -c1-300
SB2_TIMING=1 timeout 300 python scripts/r2_stages.py > gpurun_out/r2_stages.log 2>&1; grep "^rep" gpurun_out/r2_stages.log
MINE='regex:knn_|csr_|spmm_|dense_sym|tsmm_|right_mult|lincomb|residual_|f64_to|components_out|col_absmax|mu_dot|mirror_|center_gram|reduce_copies|fuzzy_|sym_|scan_|sum_f32|add_i32|decide_kernel|lm_apply|rf_|agg_|flag_nonempty|compose_|gather_kernel|comm_stats|strength_|to_fixed|iota_|fill_u8|internal_weight|comm_min|relabel_|jacobi|rr_|label_|scatter_dec|compact_flags|reset_targets'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -c 20000 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/bench_under_ncu_r2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_sweep2 -s 1 -c 1 -o gpurun_out/prof_knn_sweep2_r2 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/prof_r2.log 2>&1
ls -la gpurun_out | head -30
