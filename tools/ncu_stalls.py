#!/usr/bin/env python
"""Per-source-line stall breakdown: ncu SASS source page (csv) joined with nvdisasm -g line info.
usage: ncu_stalls.py <src.csv> <dis.txt> <mangled kernel> <file substring> <line lo> <line hi>"""
import csv, re, sys
from collections import defaultdict
src_csv, dis, kern, fsub, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(f".text.{kern}:"))
cur = ("?", 0); per = []
for l in lines[start + 1:]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l): per.append(cur)
    if len(per) >= len(data): break
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: defaultdict(float)); tot = 0
for k in range(min(len(per), len(data))):
    f, ln = per[k]; r = data[k]
    s = float(r[ci["# Samples"]] or 0); tot += s
    if fsub in f and lo <= ln <= hi:
        a = agg[ln]; a["samples"] += s; a["instr"] += float(r[ci["Instructions Executed"]] or 0)
        for h in stall_cols: a[h] += float(r[ci[h]] or 0)
tsel = sum(a["samples"] for a in agg.values())
print(f"selected range: {100*tsel/tot:.1f}% of all samples")
summ = defaultdict(float)
for ln in sorted(agg):
    a = agg[ln]
    for h in stall_cols: summ[h] += a[h]
    top = sorted(((a[h], h) for h in stall_cols), reverse=True)[:3]
    if a["samples"] > 0.002 * tot:
        print(f"{ln:5d} samples {100*a['samples']/tot:5.2f}% instr {a['instr']:.3g} | " + ", ".join(f"{h[6:]} {100*v/max(a['samples'],1):.0f}%" for v, h in top))
print("range totals:", ", ".join(f"{h[6:]} {100*v/max(tsel,1):.0f}%" for v, h in sorted(((v, h) for h, v in summ.items()), reverse=True)[:6]))
