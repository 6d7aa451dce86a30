#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
   python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_summary.csv"""
import csv, sys, collections, io
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    a = agg[r["Kernel Name"]]
    a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1")
print(f"# {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms total device time (cold-cache, serialised: compare shares)")
print("# the timed step is ONE launch of astc::encode4x4_kernel; torch kernels are the synthetic-texture generation (outside the step)")
print("launches,total_ms,avg_us,share,kernel")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{ns / 1e6:.3f},{ns / n / 1e3:.2f},{ns / tot:.4f},{k[:120]}")
