"""ncu launch list (CSV of `--metrics gpu__time_duration.sum`) -> per-kernel totals and shares (profiles/*_launch_summary_*.json).
usage: python tools/ncu_launch_summary.py launches.csv summary.json"""
import csv, json, re, sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src, newline="") if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v * {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3}[unit]
    name = re.sub(r"\(.*$", "", r["Kernel Name"])[:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
total = sum(a[1] for a in agg.values())
ours = sum(a[1] for k, a in agg.items() if "hd::" in k)
top = [{"kernel": k, "launches": a[0], "ms": a[1], "share": a[1] / total} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]]
json.dump({"total_ms": total, "ours_ms": ours, "launches": sum(a[0] for a in agg.values()),
           "note": "ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none, one eager train step "
                   "(tools/profile_step.py, TF32 matmul as bench.py); serialised, cold-cache per-launch times",
           "top": top}, open(dst, "w"), indent=1)
print("total ms", total, "ours ms", ours, "launches", sum(a[0] for a in agg.values()))
