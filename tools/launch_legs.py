"""Splits an ncu launch list of bench.py (tools/launch_summary.py input) into the legs of the bench: device-resident calls
(classify + whole-grid flux), zero-copy host-array calls (flux alone) and staged jt==1 host-array calls (chunks)."""
import csv, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = [r for r in csv.DictReader(rows) if r["Metric Name"] == "gpu__time_duration.sum"]


def us(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v / 1e3 if r["Metric Unit"] == "ns" else v * 1e3 if r["Metric Unit"] == "ms" else v


grid = lambda r: int(r["Grid Size"].strip("()").split(",")[0])
whole = max(grid(r) for r in rd if r["Kernel Name"].startswith("classify"))
legs = {"device-resident (classify + whole-grid flux)": ([], []), "zero-copy host-array calls (flux on pinned host memory, PCIe-bound)": ([], []),
        "staged jt==1 host-array calls (chunks)": ([], [])}
prev = None
for r in rd:
    if r["Kernel Name"].startswith("void flux_kernel"):
        if prev is not None and prev["Kernel Name"].startswith("classify"):
            k = "device-resident (classify + whole-grid flux)" if grid(prev) == whole else "staged jt==1 host-array calls (chunks)"
            legs[k][0].append(us(r)); legs[k][1].append(us(prev))
        else:
            legs["zero-copy host-array calls (flux on pinned host memory, PCIe-bound)"][0].append(us(r))
    prev = r
print("by leg of bench.py:")
for k, (f, c) in legs.items():
    if f:
        tf, tc = sum(f), sum(c)
        extra = f"; classify mean {tc / len(c):.1f} us = {100 * tc / (tc + tf):.1f} % of the leg, flux {100 * tf / (tc + tf):.1f} %" if c else ""
        print(f"  {k}: {len(f)} flux launches, mean {tf / len(f):.1f} us{extra}")
