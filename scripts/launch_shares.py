"""Per-kernel launch counts / time shares from an ncu launch list (gpu__time_duration.sum --csv).
usage: python scripts/launch_shares.py gpurun_out/launches_<tag>.csv > profiles/<tag>_launch_shares.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.OrderedDict()
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("tvl1::", "")
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(u, 1e-6)
    a = tot.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
s = sum(v for _, v in tot.values()) or 1.0
print("kernel,launches,total_ms,share")
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.3f,%.3f" % (k, n, v, v / s))
