#!/bin/bash
# ncu --set full of Leiden's decide kernels (local moving <0> and refinement <1>, first level-0 launches) at the bench workload
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none -k regex:decide_kernel -s 2 -c 4 -o gpurun_out/prof_decide_r2 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/prof_decide_r2.log 2>&1
timeout 120 ncu -i gpurun_out/prof_decide_r2.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/prof_decide_r2_raw.csv
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/prof_decide_r2_raw.csv')))
h,u=rows[0],rows[1]
want=["Kernel Name","gpu__time_duration.sum","launch__grid_size","launch__registers_per_thread","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sector_hit_rate.pct","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum"]
for r in rows[2:]:
    print("---")
    for i,k in enumerate(h):
        if k in want or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(r[i] or 0)>0.5):
            print(f"{k} [{u[i]}] = {r[i]}")
P
