#!/bin/bash
# ncu evidence for round 1 (run under gpurun, 1 GPU).  Numbers printed by runs under ncu are NOT bench values.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
# (1) launch list of one full step at the bench workload
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
# (2) full capture of the dominant kernel at a quarter-size workload (40 replays of a 0.25 s launch)
ncu --set full --clock-control none --import-source on -k regex:knn_pass1 -c 1 -o gpurun_out/prof_knn_pass1_r1 \
    python bench.py --n-cells 325000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/prof1.log 2>&1
# (3) full capture of the HBM/L2-bound kernels (one launch each) at 325k
ncu --set full --clock-control none --import-source on -k regex:'csr_gram_kernel|spmm_csr_kernel|csr_col_stats|fuzzy_rows|sym_count|knn_rescore|knn_prep' -c 7 \
    -o gpurun_out/prof_hbm_kernels_r1 python bench.py --n-cells 325000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/prof2.log 2>&1
ncu --set full --clock-control none -k regex:'decide_kernel|agg_insert' -c 3 \
    -o gpurun_out/prof_leiden_r1 python bench.py --n-cells 325000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/prof3.log 2>&1
ls -la gpurun_out
