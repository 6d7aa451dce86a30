#!/usr/bin/env python3
"""Static register-file read count of a SASS region (the operand-delivery model of DESIGN.md 4.1b).
   cuobjdump -sass -fun <name> file | python tools/sass_reads.py [--clock]
With --clock the region between the two `CS2R ..., SR_CLOCKLO` reads (a microbenchmark's timed loop) is
analysed; otherwise the whole listing.  Cost per instruction: FMA-pipe packed ops max(2, reads/2), scalar
FMA-pipe ops max(1, reads/2), everything else reads/2 (issue slot >= 1 is reported separately)."""
import re, sys, collections

reg_re = re.compile(r"^[-|~!]*R(\d+)((?:\.[A-Za-z0-9_]+)*)\|?$")
addr_re = re.compile(r"\[(?:R(\d+)(\.64|\.U32|\.X\d+)*)?([^\]]*)\]")
NO_DEST = ("ST", "STG", "STS", "STL", "RED", "BAR", "BRA", "EXIT", "BSYNC", "BSSY", "NOP", "LDGSTS", "LDGDEPBAR", "DEPBAR", "WARPSYNC", "CALL", "RET")
PACKED = ("FFMA2", "FMUL2", "FADD2")
SCALAR_FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2")

lines = []
for l in sys.stdin:
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*;", l)
    if m: lines.append(m.group(1))
if "--clock" in sys.argv:
    idx = [i for i, l in enumerate(lines) if "SR_CLOCKLO" in l]
    lines = lines[idx[0] + 1: idx[1]]
prev = {}
tot = collections.Counter(); n = collections.Counter(); cost = 0.0; reuse_hits = 0
for src in lines:
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)\s*(.*)$", src)
    if not m: continue
    op, mods, rest = m.group(2), m.group(3) or "", m.group(4)
    toks = [t.strip() for t in re.split(r",(?![^\[]*\])", rest) if t.strip()]
    srcs = toks if op in NO_DEST else toks[1:]
    reads = 0; slot = 0; cur = {}
    for t in srcs:
        if re.match(r"^!?U?P(T|\d+)$", t): continue
        am = addr_re.search(t)
        if am:
            if am.group(1) is not None: reads += 1
            slot += 1; continue
        rm = reg_re.match(t)
        if rm:
            reg = int(rm.group(1)); suf = rm.group(2) or ""
            width = 2 if (".F32x2" in suf or ".64" in suf) else 1
            if op.startswith("ST") and "128" in mods: width = 4
            if prev.get(slot) == reg: reuse_hits += 1
            else: reads += width
            if ".reuse" in suf: cur[slot] = reg
        slot += 1
    prev = cur
    tot[op] += reads; n[op] += 1
    cost += max(2.0, reads / 2) if op in PACKED else max(1.0, reads / 2) if op in SCALAR_FMA else max(1.0, reads / 2)
print(f"{sum(n.values())} instructions, {sum(tot.values())} RF reads ({reuse_hits} operands from the reuse cache) -> {sum(tot.values()) / 2:.0f} cycles at 2 reads/clk; per-instruction max(pipe, reads/2) model: {cost:.0f} cycles")
print("  " + "  ".join(f"{op}:{n[op]}x{tot[op] / n[op]:.2f}" for op, _ in n.most_common(10)))
