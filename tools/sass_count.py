#!/usr/bin/env python
"""Static SASS mnemonic counts of dmb_solve_kernel in generated libraries: python tools/sass_count.py name:<kcache hash> ...  (profiles/r2_sass_evidence.json)"""
import subprocess, re, sys, json
libs = dict(a.split(":") for a in sys.argv[1:])
pats = ["DMMA.8x8x4", "UBLKCP.S.G", "UBLKCP.G.S", "SYNCS", "CREDUX", "DFMA", "DADD", "DMUL", "MUFU.RCP64H", "SHFL", "BAR.SYNC", "LDS", "STS", "LDG", "STG", "LDL", "STL", "UTC"]
out = {}
for name, h in libs.items():
    txt = subprocess.run(["cuobjdump", "-sass", f"/root/repo/do_mpc_b200/_kcache/{h}/libdmb_{h}.so"], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)
    for p in parts[1:]:
        fn = p.split("\n", 1)[0].strip()
        if "dmb_solve_kernel" not in fn: continue
        ins = re.findall(r"^\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", p, flags=re.M)
        d = {"sass_instructions": len(ins)}
        for pt in pats:
            d[pt] = sum(1 for i in ins if re.search(r"(^|\s)" + re.escape(pt), i))
        out[name] = d
print(json.dumps(out, indent=1))
