#!/usr/bin/env python3
"""Per-phase breakdown of an encode kernel from an ncu source page: instructions, RF operand
reads and stall samples by reason, cut at landmark instructions of the hot loop.
   python tools/ncu_phases.py prof.ncu-rep warps"""
import csv, io, re, subprocess, sys, collections

rep, warps = sys.argv[1], int(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO("\n".join(txt.splitlines()[1:]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
NO_DEST = ("ST", "STG", "STS", "STL", "RED", "BAR", "BRA", "EXIT", "BSYNC", "BSSY", "NOP", "LDGSTS", "LDGDEPBAR", "DEPBAR", "WARPSYNC", "CALL", "RET")
reg_re = re.compile(r"^[-|~!]*R(\d+)((?:\.[A-Za-z0-9_]+)*)\|?$")
addr_re = re.compile(r"\[(?:R(\d+)(\.64|\.U32|\.X\d+)*)?([^\]]*)\]")
inst = []
prev = {}
for r in rows[1:]:
    if len(r) < len(hdr): continue
    src = r[col["Source"]].strip(); n = int(r[col["Instructions Executed"]] or 0) / warps
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)\s*(.*?)\s*;?$", src)
    if not m: continue
    op, mods, rest = m.group(2), m.group(3) or "", m.group(4)
    toks = [t.strip() for t in re.split(r",(?![^\[]*\])", rest) if t.strip()]
    srcs = toks if op in NO_DEST else toks[1:]
    reads = 0; slot = 0; cur = {}
    for t in srcs:
        if re.match(r"^!?U?P(T|\d+)$", t): continue
        am = addr_re.search(t)
        if am:
            if am.group(1) is not None: reads += 2 if (am.group(2) or "").startswith(".64") else 1
            slot += 1; continue
        rm = reg_re.match(t)
        if rm:
            reg = int(rm.group(1)); suf = rm.group(2) or ""
            width = 2 if (".F32x2" in suf or ".64" in suf) else 1
            if op.startswith("ST") and "128" in mods: width = 4
            elif op.startswith("ST") and "64" in mods: width = 2
            if prev.get(slot) != reg: reads += width
            if ".reuse" in suf: cur[slot] = reg
        slot += 1
    prev = cur
    st = {s: int(r[col[s]] or 0) for s in stall_cols}
    inst.append(dict(op=op, n=n, reads=reads, samples=int(r[col["# Samples"]] or 0), st=st, src=src))
hot = [i for i in inst if i["n"] >= 0.5]
# landmarks: first MUFU.RSQ (PI starts ~40 before), last MUFU.RSQ/RCP pair of the PI, first FMNMX, LDS run, STG
def first(pred, start=0):
    for k in range(start, len(hot)):
        if pred(hot[k]): return k
    return len(hot)
k_ldg = first(lambda i: i["op"] == "LDG")
k_rsq = first(lambda i: i["src"].lstrip().startswith("MUFU.RSQ"))
rsqs = [k for k, i in enumerate(hot) if "MUFU.RSQ" in i["src"] and not i["src"].startswith("@")]
k_pi_end = rsqs[-1] + 8 if rsqs else k_rsq
k_lds = first(lambda i: i["op"] == "LDS", k_pi_end)
k_prmt_end = max([k for k, i in enumerate(hot[:k_rsq]) if i["op"] == "PRMT"] + [0])
cuts = [("convert", 0, k_prmt_end + 8), ("loop-ctl/cov", k_prmt_end + 8, k_rsq - 10), ("power-iter", k_rsq - 10, k_pi_end),
        ("minmax/weights", k_pi_end, k_lds - 30), ("quant/pack/store", k_lds - 30, len(hot))]
tot_s = sum(i["samples"] for i in inst)
print(f"warp-instr per warp-block {sum(i['n'] for i in inst):.1f}; RF reads {sum(i['n'] * i['reads'] for i in inst):.1f}; samples {tot_s}")
print(f"{'phase':18s} {'instr':>7s} {'reads':>7s} {'%time':>6s}  top stall reasons (% of phase samples)")
for name, a, b in cuts:
    s = hot[a:b]
    smp = sum(i["samples"] for i in s)
    agg = collections.Counter()
    for i in s:
        for kk, v in i["st"].items(): agg[kk] += v
    tops = ", ".join(f"{k[6:]} {100 * v / max(1, sum(agg.values())):.0f}" for k, v in agg.most_common(5))
    print(f"{name:18s} {sum(i['n'] for i in s):7.1f} {sum(i['n'] * i['reads'] for i in s):7.1f} {100 * smp / tot_s:6.1f}  {tops}")
