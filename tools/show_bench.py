#!/usr/bin/env python
"""print the interesting parts of a bench.py JSON line: python tools/show_bench.py gpurun_out/x.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "gpu_launches", "clocks")})
print("e2e", d.get("e2e"))
r = d.get("roofline", {})
print("roofline", {k: r.get(k) for k in ("achieved", "frac", "traffic", "traffic_source", "kernel_ms")})
print("config", d.get("config"))
print("cpu_baseline", d.get("cpu_baseline"))
print("strong", d.get("strong"))
for k, v in (d.get("other_configs") or {}).items():
    print(k, {a: v[a] for a in ("solves_per_s", "success_rate", "iters_mean", "ms_per_step")}, "frac %.4f" % v["roofline"]["frac"], v["clocks"])
