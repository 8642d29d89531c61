"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total time, share, count.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [top]"""
import csv, sys, re, collections
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"]
    name = re.sub(r"\(anonymous namespace\)", "<unnamed>", name)
    t = tot.setdefault(name, [0.0, 0])
    t[0] += us; t[1] += 1
total = sum(t[0] for t in tot.values()); n = sum(t[1] for t in tot.values())
print("total: %d kernel launches, %.1f ms summed device time" % (n, total / 1e3))
print("%10s %6s %6s  %s" % ("us", "share", "count", "kernel"))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for name, (us, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%10.0f %5.1f%% %6d  %s" % (us, 100 * us / total, c, name[:120]))
