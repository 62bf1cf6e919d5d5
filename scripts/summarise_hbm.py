"""Per-kernel duration + DRAM bytes table from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`.
usage: python scripts/summarise_hbm.py in.csv [peak_gbs]"""
import csv, re, sys
from collections import defaultdict
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6587.7
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
per = defaultdict(dict)
names = {}
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    mult = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per[r["ID"]][r["Metric Name"]] = v * mult
    n = re.sub(r"^void ", "", r["Kernel Name"]); n = re.sub(r"\(.*$", "", n).replace("<unnamed>::", "")
    names[r["ID"]] = n
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, m in per.items():
    a = agg[names[i]]
    a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0.0); a[2] += m.get("dram__bytes_read.sum", 0.0); a[3] += m.get("dram__bytes_write.sum", 0.0)
print("# HBM/L2-bound kernels at the bench workload (1.3M x 2000, n_pcs 50, k 15), one bench step under ncu (cold, serialised launches)")
print(f"# achieved = (dram read + write) / duration; frac of the measured HBM peak {peak} GB/s (MEASURED_PEAKS.json hbm_gbs)")
print("kernel,launches,total_ms,dram_read_MB,dram_write_MB,achieved_GBs,frac_of_hbm_peak")
for n, (c, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
    print(f"{n},{c},{t*1e3:.3f},{rd/1e6:.1f},{wr/1e6:.1f},{gbs:.0f},{gbs/peak:.3f}")
