#!/usr/bin/env python
"""Print a bench JSON line's headline and the per-kernel table written by `bench.py --profile-out`.
usage: show_kernels.py bench.json kernels.json [hbm|tensor|all] [min_ms]"""
import json, sys
b = json.load(open(sys.argv[1])); k = json.load(open(sys.argv[2]))["kernels"]
kind = sys.argv[3] if len(sys.argv) > 3 else "all"
min_ms = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
r = b["roofline"]
print(f"{b['value']:.1f} img/s  {b['ms_per_step']:.2f} ms/step  e2e {b['e2e']['value']:.1f}  tensor {r['all_tcgen05']['ms_per_step']:.2f} ms @ {r['all_tcgen05']['achieved']:.0f} TF/s (conv {r['achieved']:.0f})  "
      f"hbm {r['hbm_kernels']['ms_per_step']:.2f} ms @ {r['hbm_kernels']['achieved_gbs']:.0f} GB/s  clocks {b['clocks']}")
for n, v in sorted(k.items(), key=lambda kv: -kv[1]["ms"]):
    if (kind == "all" or v["kind"] == kind) and v["ms"] >= min_ms:
        print(f"{n:40s} {v['kind']:6s} n={v['launches']:3d} ms={v['ms']:7.3f} {v['achieved']:8.1f} {v['unit']}")
