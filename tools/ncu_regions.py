#!/usr/bin/env python3
"""Stall samples of a kernel by address window (N equal windows over the hot instructions), plus
shared-memory wavefront totals: where in the instruction stream the time goes.
   python tools/ncu_regions.py prof.ncu-rep [windows]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
nwin = int(sys.argv[2]) if len(sys.argv) > 2 else 24
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO("\n".join(txt.splitlines()[1:]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
inst = []
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    g = lambda name: int(float(r[col[name]] or 0)) if name in col and r[col[name]] not in ("", "-") else 0
    inst.append(dict(src=r[col["Source"]].strip(), n=g("Instructions Executed"), samples=g("# Samples"),
                     st={s: g(s) for s in stall_cols}, wf=g("L1 Wavefronts Shared"), wfi=g("L1 Wavefronts Shared Ideal")))
tot = sum(i["samples"] for i in inst)
nmax = max(i["n"] for i in inst)
print(f"instructions {len(inst)}, samples {tot}, warp-level executions of the hottest instruction {nmax}")
print(f"dynamic warp-instructions {sum(i['n'] for i in inst)}  ({sum(i['n'] for i in inst) / nmax:.1f} per hot-loop trip)")
print(f"shared wavefronts {sum(i['wf'] for i in inst)} (ideal {sum(i['wfi'] for i in inst)}) = {sum(i['wf'] for i in inst) / nmax:.1f} per trip")
agg = collections.Counter()
for i in inst:
    for s, v in i["st"].items():
        agg[s] += v
print("stall totals:", ", ".join(f"{k[6:]} {v * 100 / tot:.1f}%" for k, v in agg.most_common(9)))
per = (len(inst) + nwin - 1) // nwin
for w in range(nwin):
    seg = inst[w * per:(w + 1) * per]
    if not seg:
        break
    s = sum(i["samples"] for i in seg)
    c = collections.Counter()
    for i in seg:
        for k, v in i["st"].items():
            c[k] += v
    ops = collections.Counter(i["src"].split()[0 if not i["src"].startswith("@") else 1].split(".")[0] for i in seg if i["n"] > nmax * 0.3)
    dyn = sum(i["n"] for i in seg) / nmax
    print(f"[{w * per:5d}..] {s * 100 / tot:5.1f}% samples, {dyn:6.1f} instr/trip | " + ", ".join(f"{k[6:]} {v * 100 / max(s, 1):.0f}%" for k, v in c.most_common(4))
          + " | " + " ".join(f"{k}:{v}" for k, v in ops.most_common(5)))
