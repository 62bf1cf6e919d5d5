"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python scripts/summarise_launches.py in.csv [title]"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("<unnamed>::", "")
    ns = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] == "us": ns *= 1e3
    if r["Metric Unit"] == "ms": ns *= 1e6
    rows.append((name, ns))
tot = sum(ns for _, ns in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1; agg[n][1] += ns
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes, with bench.py's live numbers")
print(f"# total {tot/1e6:.1f} ms in {len(rows)} launches")
print("kernel,launches,total_ms,share")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{c},{ns/1e6:.3f},{100*ns/tot:.2f}%")
