#!/bin/bash
# measurement of the widened rows' kernels: duration + DRAM bytes (ncu) while r2_widened_perf.py runs at 1.3M cells
# pass 1 (done, profiles/r2_widened_kernels_eigs.txt): -k "regex:eig_" -c 700; pass 2: the UMAP epochs and the small streaming kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:umap_epoch|umap_eps|umap_wmax|umap_minmax|umap_rescale|group_arc|col_stats_rows|scale_cols" -c 240 --csv --log-file gpurun_out/widened_kernels_r2b.csv \
    python scripts/r2_widened_perf.py 1300000 > gpurun_out/widened_under_ncu_b.log 2>&1
tail -3 gpurun_out/widened_under_ncu_b.log
python scripts/summarise_hbm.py gpurun_out/widened_kernels_r2b.csv | head -30
timeout 300 ncu --set full --clock-control none -k regex:umap_epoch -s 50 -c 1 -o gpurun_out/prof_umap_epoch python scripts/r2_widened_perf.py 400000 > gpurun_out/prof_umap.log 2>&1
timeout 120 ncu -i gpurun_out/prof_umap_epoch.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for i,k in enumerate(h):
    if k in ('gpu__time_duration.sum','dram__bytes_read.sum','lts__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread') or (k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') and float(v[i] or 0)>0.3):
        print(k,v[i],u[i])
"
