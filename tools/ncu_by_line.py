#!/usr/bin/env python
"""Aggregate an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) by CUDA source line, using
`nvdisasm -g -c` line annotations of the same cubin (instructions are matched by address order).

usage: ncu_by_line.py <src.csv from ncu> <nvdisasm output> <kernel mangled name> [top N]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, dis, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    lines = open(dis).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(f".text.{kern}:"))
    cur = ("?", 0)
    per_instr = []
    for l in lines[start + 1:]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
            per_instr.append(cur)
        if l.startswith(".text.") or l.startswith("\t.section") and per_instr:
            if len(per_instr) >= len(data):
                break
    n = min(len(per_instr), len(data))
    agg = defaultdict(lambda: [0, 0, 0.0])
    tot_i = tot_s = 0
    for k in range(n):
        r = data[k]
        ie = int(float(r[ci["Instructions Executed"]] or 0))
        ss = int(float(r[ci["# Samples"]] or 0))
        a = agg[per_instr[k]]
        a[0] += ie; a[1] += ss
        tot_i += ie; tot_s += ss
    print(f"instructions matched {n} (ncu {len(data)}, disasm {len(per_instr)}); total warp-instr {tot_i:.3e}, samples {tot_s}")
    byfile = defaultdict(lambda: [0, 0])
    for (f, ln), (ie, ss, _) in agg.items():
        byfile[f][0] += ie; byfile[f][1] += ss
    for f, (ie, ss) in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
        print(f"  {f:24s} instr {100*ie/tot_i:5.1f}%  samples {100*ss/max(tot_s,1):5.1f}%")
    print("top lines by stall samples:")
    for (f, ln), (ie, ss, _) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {f}:{ln:<5d} instr {100*ie/tot_i:5.2f}%  samples {100*ss/max(tot_s,1):5.2f}%")


if __name__ == "__main__":
    main()
