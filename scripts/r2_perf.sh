#!/bin/bash
# Round-2 perf pass: Gram A/B, bench, stage times, ncu of the kNN sweep (DRAM traffic after the mean-front change), launch list.
mkdir -p gpurun_out
timeout 200 python scripts/r2_gram.py 100000 2>&1 | grep -E "gram|max rel" 
SB2_GRAM_V1=1 timeout 200 python scripts/r2_gram.py 100000 2>&1 | grep -E "gram|max rel"
timeout 200 python scripts/r2_gram.py 2>&1 | grep -E "gram"
SB2_GRAM_V1=1 timeout 200 python scripts/r2_gram.py 2>&1 | grep -E "gram"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',{k:d['e2e'][k] for k in ('value','s_per_step')}); print('roofline',{k:d['roofline'].get(k) for k in ('achieved','frac','launch_ms','frac_issued','share_of_step')}); print('stages',d['stages']); print('cpu',d.get('cpu_baseline',{}).get('value')); print(d.get('clocks'))
P
tail -3 gpurun_out/r2b_bench.err | cut -c1-300
SB2_TIMING=1 timeout 300 python scripts/r2_stages.py > gpurun_out/r2b_stages.log 2>&1; grep "^rep" gpurun_out/r2b_stages.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_sweep2 -s 1 -c 1 -o gpurun_out/prof_knn_sweep2_r2b \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/prof_r2b.log 2>&1
timeout 300 ncu -i gpurun_out/prof_knn_sweep2_r2b.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for k in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):
    i=h.index(k); print(k,v[i],u[i])
"
MINE='regex:knn_|csr_|spmm_|dense_sym|tsmm_|right_mult|lincomb|residual_|f64_to|components_out|col_absmax|mu_dot|mirror_|center_gram|reduce_copies|fuzzy_|sym_|scan_|sum_f32|add_i32|decide_kernel|lm_apply|rf_|agg_|flag_nonempty|compose_|gather_kernel|comm_stats|strength_|to_fixed|iota_|fill_u8|internal_weight|comm_min|relabel_|jacobi|rr_|label_|scatter_dec|compact_flags|reset_targets|gram_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -c 20000 --csv --log-file gpurun_out/launches_r2b.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/bench_under_ncu_r2b.log 2>&1
ls -la gpurun_out | tail -8
