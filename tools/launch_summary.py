"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total, mean, share."""
import csv, sys, collections
path = sys.argv[1]
rows = [l for l in open(path) if l.startswith('"')]
rd = csv.DictReader(rows)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] == "ns":
        v /= 1e3
    elif r["Metric Unit"] == "ms":
        v *= 1e3
    k = r["Kernel Name"]
    tot[k][0] += 1
    tot[k][1] += v
allus = sum(v[1] for v in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'mean_us':>10s} {'share':>7s}")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / allus:6.1f}%")
