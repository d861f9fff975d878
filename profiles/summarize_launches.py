"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); total = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    agg[name][0] += 1; agg[name][1] += v; total += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t/1e3:9.3f} ms {100*t/total:5.1f}% n={n:4d} avg={t/n:9.1f} us  {k[:100]}")
print(f"total {total/1e3:.3f} ms")
