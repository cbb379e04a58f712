"""profiles/traffic.json from the raw ncu metric pages of tools/ncu_all.sh (one flux-kernel launch per variant of the
CURRENT build): per kernel variant, per grid point --
  fp64_thread_inst_per_point   FP64-pipe warp instructions executed x 32 / points (pipe slots, idle lanes included:
                               what occupies the pipe), the numerator of bench.py's roofline.frac
  dram_bytes_per_point         dram__bytes_read.sum + dram__bytes_write.sum per point (`roofline.traffic`)
  pipe_fp64_pct, issue_pct, lanes_per_inst, registers, spill_inst_per_point, duration_us   for the record
usage: python tools/make_traffic.py <tag> [Ni Nj]   (reads gpurun_out/ncu_full_<tag>_<case>_raw.csv)"""
import csv, glob, json, os, sys

tag = sys.argv[1]
Ni, Nj = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1440, 720)
n = Ni * Nj
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"andreas@nb30": "andreas@nb30", "coare3p0@nb30": "coare3p0@nb30", "ncar": "ncar", "andreas": "andreas", "coare3p0": "coare3p0", "coare3p6": "coare3p6", "ecmwf": "ecmwf",
         "coare3p6_skin_day": "coare3p6+skin", "ecmwf_skin_day": "ecmwf+skin", "coare3p0_skin_day": "coare3p0+skin",
         "coare3p6_skin_night": "coare3p6+skin@night", "ecmwf_skin_night": "ecmwf+skin@night",
         "andreas_nb30": "andreas@nb30", "coare3p0_nb30": "coare3p0@nb30"}
out = {}
for f in sorted(glob.glob(os.path.join(root, "gpurun_out", f"ncu_full_{tag}_*_raw.csv"))):
    case = os.path.basename(f)[len(f"ncu_full_{tag}_"):-len("_raw.csv")]
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        continue
    d = dict(zip(rows[0], rows[2]))
    units = dict(zip(rows[0], rows[1]))
    SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ms": 1e3, "ns": 1e-3, "s": 1e6}   # bytes / microseconds
    g = lambda k: float(d[k].replace(",", "")) * SCALE.get(units.get(k, ""), 1.0) if d.get(k) not in (None, "", "n/a") else None
    warp64 = g("sm__inst_executed_pipe_fp64.sum")
    if warp64 is None:
        # `--set full` reports the pipe as a percentage only: 2 FP64 warp instructions per SM cycle at 100 % (64 lanes)
        pct, cyc = g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), g("sm__cycles_active.sum")
        warp64 = pct / 100.0 * 2.0 * cyc if (pct and cyc) else None
    out[NAMES.get(case, case)] = {
        "kernel": d.get("Kernel Name"), "grid": [Ni, Nj], "source": f"ncu --set full, {os.path.basename(f)}",
        "fp64_thread_inst_per_point": warp64 * 32 / n if warp64 else None,
        "dram_bytes_per_point": ((g("dram__bytes_read.sum") or 0) + (g("dram__bytes_write.sum") or 0)) / n,
        "pipe_fp64_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lanes_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "inst_per_point": (g("smsp__inst_executed.sum") or 0) * 32 / n,
        "registers": g("launch__registers_per_thread"),
        "spill_inst_per_point": (g("sass__inst_executed_register_spilling") or 0) * 32 / n,
        "duration_us": g("gpu__time_duration.sum"),
    }
for a in ("andreas", "coare3p0"):
    if a in out and f"{a}@nb30" in out and out[a]["fp64_thread_inst_per_point"] and out[f"{a}@nb30"]["fp64_thread_inst_per_point"]:
        out[a]["fp64_thread_inst_per_iteration"] = (out[f"{a}@nb30"]["fp64_thread_inst_per_point"] - out[a]["fp64_thread_inst_per_point"]) / 25.0
# C2: the 24 launches of one session (night and day launches execute different amounts of work): mean per launch
c2 = os.path.join(root, "gpurun_out", f"ncu_metrics_{tag}_c2_session.csv")
if os.path.exists(c2):
    rows = [r for r in csv.reader(open(c2)) if len(r) > 10]
    hdr = rows[0]
    iN, iV, iU = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    vals = {}
    for r in rows[1:]:
        v = float(r[iV].replace(",", ""))
        if r[iN] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iU], 1.0)
        vals.setdefault(r[iN], []).append(v)
    k64 = vals.get("sm__inst_executed_pipe_fp64.sum", [])
    if k64:
        out["coare3p6+skin@C2session"] = {
            "kernel": "flux_kernel<COARE3P6,skin,zt!=zu>, the 24 launches of one BASELINE C2 session", "grid": [Ni, Nj],
            "source": f"ncu --metrics, {os.path.basename(c2)}", "launches": len(k64),
            "fp64_thread_inst_per_point": sum(k64) / len(k64) * 32 / n,
            "fp64_thread_inst_per_point_min_max": [min(k64) * 32 / n, max(k64) * 32 / n],
            "pipe_fp64_pct": sum(vals["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]) / len(k64),
            "issue_pct": sum(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / len(k64),
            "lanes_per_inst": sum(vals["smsp__thread_inst_executed_per_inst_executed.ratio"]) / len(k64),
            "duration_us": sum(vals["gpu__time_duration.sum"]) / len(k64), "dram_bytes_per_point": None, "registers": None,
            "inst_per_point": sum(vals["smsp__inst_executed.sum"]) / len(k64) * 32 / n, "spill_inst_per_point": 0.0}
json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
for k, v in out.items():
    print(f"{k:24s} {v['duration_us']:8.1f} us  fp64/pt {v['fp64_thread_inst_per_point']:8.0f}  pipe {v['pipe_fp64_pct']:5.1f} %  issue {v['issue_pct']:5.1f} %  "
          f"lanes {v['lanes_per_inst']:5.2f}  dram B/pt {v['dram_bytes_per_point'] or 0:6.1f}  spills/pt {v['spill_inst_per_point']:5.1f}")
