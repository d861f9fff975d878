"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python profiles/launch_summary.py
<csv> [steps]  -> microseconds per step, launches per step and share of the summed device time."""
import collections
import csv
import sys

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr, agg = None, collections.OrderedDict()
for r in csv.reader(open(path)):
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(d["Kernel Name"][:110], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"sum of kernel durations: {tot / 1000 / steps:.1f} us per step ({steps} steps)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
    print(f"{a[1] / 1000 / steps:10.1f} us {a[0] / steps:7.1f} x {100 * a[1] / tot:5.1f}%  {k}")
