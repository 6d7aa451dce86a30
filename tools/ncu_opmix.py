#!/usr/bin/env python3
"""Dynamic per-opcode instruction mix + top stall sites from an ncu source page.
   python tools/ncu_opmix.py prof.ncu-rep [warps]"""
import csv, io, subprocess, sys, re, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
# first line = kernel name
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
mix = collections.Counter(); samp = collections.Counter(); total = 0; tsamp = 0
sites = []
for r in rows[1:]:
    if len(r) <= ie: continue
    src = r[ia].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else src.split()[0]
    n = int(r[ie] or 0); s = int(r[isamp] or 0)
    mix[op] += n; samp[op] += s; total += n; tsamp += s
    sites.append((s, n, src))
warps = int(sys.argv[2]) if len(sys.argv) > 2 else max(n for _, n, _ in sites)
print(f"total warp-instructions {total}  per warp {total / warps:.1f}  (warps {warps})")
print(f"{'op':10s} {'per-warp':>9s} {'%inst':>6s} {'%samples':>8s}")
for op, n in mix.most_common(28):
    print(f"{op:10s} {n / warps:9.1f} {100 * n / total:6.1f} {100 * samp[op] / max(1, tsamp):8.1f}")
print("-- top stall sites")
for s, n, src in sorted(sites, reverse=True)[:25]:
    print(f"{100 * s / max(1, tsamp):5.2f}%  x{n / warps:5.2f}  {src[:90]}")
