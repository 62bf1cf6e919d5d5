#!/bin/bash
# ncu evidence, round 1 final kernels (tensor-core kNN).  Numbers printed under ncu are NOT bench values.
set -x
mkdir -p gpurun_out
MINE='regex:knn_|csr_|spmm_|dense_sym|tsmm_|right_mult|lincomb|residual_|f64_to|components_out|col_absmax|mu_dot|mirror_|center_gram|reduce_copies|fuzzy_|sym_|scan_|sum_f32|add_i32|decide_kernel|lm_apply|rf_|agg_|flag_nonempty|compose_|gather_kernel|comm_stats|strength_|to_fixed|iota_|fill_u8|internal_weight|comm_min|relabel_'
# (1) launch list of one full step at the bench workload (this repo's kernels only; torch's data generation excluded)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -c 20000 --csv --log-file gpurun_out/launches_r1c.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_c.log 2>&1
# (2) full capture of the dominant kernel at the bench workload (1.3M: one launch, ~40 replays of 0.45 s)
ncu --set full --clock-control none --import-source on -k regex:knn_pass1_tc -s 1 -c 1 -o gpurun_out/prof_knn_tc_r1c \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/prof1c.log 2>&1
ls -la gpurun_out
