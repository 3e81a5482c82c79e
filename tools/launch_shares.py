#!/usr/bin/env python
"""Per-kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_shares.py
launches.csv out.json ["source note"]. Per-launch times under ncu are cold-cache and serialised: compare SHARES."""
import collections, csv, json, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, total, n = collections.OrderedDict(), 0.0, 0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    us = float(r[mv].replace(",", "")) / (1000.0 if r[mu] == "ns" else 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
    n += 1
out = {"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1], "launches": n, "total_us": total,
       "kernels": {k: {"launches": a[0], "us": round(a[1], 1), "share": round(a[1] / total, 4)}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
tc = sum(v["share"] for k, v in out["kernels"].items() if "_tc_kernel" in k or "wgrad_halo" in k)
out["tcgen05_share"] = round(tc, 4)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(f"{n} launches, {total:.0f} us, tcgen05 share {tc:.3f}")
