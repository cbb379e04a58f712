#!/usr/bin/env python
"""bench.py -- headline benchmark of the aerobulk_model hot path on B200.

Metric (BASELINE.json): grid points / s (FP64).  One STEP = one 24-hour session of
BASELINE config 2: COARE 3.6 with cool-skin + warm-layer on a 0.25 deg 1440x720 grid,
24 hourly `aerobulk_model` calls (jt = 1..24) with the warm-layer state carried on the
device.  At N GPUs each rank owns a 1440x720 latitude row block of a 1440x(720 N) grid
(weak scaling, no data-path collective; one 64-double NCCL all-reduce per session for
AEROBULK_INIT's global field statistics).

    python bench.py [--gpus N --steps K --warmup W]           our CUDA path
    python bench.py --impl reference [...]                      the CPU oracle on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...       one rank per GPU

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every key).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO, SKIN, NI, NJ, NT, NB_ITER, ZT, ZU = "coare3p6", True, 1440, 720, 24, 5, 2.0, 10.0
IN_KEYS = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
OUT_KEYS = ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")
WORKLOAD = "C2: COARE 3.6 + cool-skin/warm-layer, 1440x720 per GPU, 24 hourly steps, nb_iter=5, zt=2 zu=10"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [s for t, s in self.samples if t0 <= t <= t1] or [s for _, s in self.samples[-5:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            c = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(threads: int, target_seconds: float = 12.0) -> dict:
    """The CPU oracle (C port of the reference algorithm, gcc -O2 -ffp-contract=off) on the host cores,
    on a bounded row sample of the same workload: 1440 x nj rows x 24 hourly steps."""
    from aerobulk_b200 import synth
    from oracle.oracle import OracleSession

    def run(nj):
        f = synth.fields(NI, NJ, j0=(NJ - nj) // 2, j1=(NJ - nj) // 2 + nj)
        rsw = [synth.rad_sw_hour(NI, NJ, jt, j0=(NJ - nj) // 2, j1=(NJ - nj) // 2 + nj) for jt in range(1, NT + 1)]
        s = OracleSession(threads=threads)
        t0 = time.perf_counter()
        for jt in range(1, NT + 1):
            s.model(jt, NT, ALGO, ZT, ZU, *[f[k] for k in IN_KEYS], Niter=NB_ITER, l_use_skin=True,
                    rad_sw=rsw[jt - 1], rad_lw=f["rad_lw"])
        return time.perf_counter() - t0

    probe_nj = 8
    run(probe_nj)                       # thread pool + page warm-up
    t = run(probe_nj)
    rate = NI * probe_nj * NT / t
    nj = int(min(NJ, max(probe_nj, round(rate * target_seconds / (NI * NT)))))
    t = run(nj)
    # the reference itself is serial: one thread on a 16-row sample, for the record
    threads_all, threads = threads, 1
    run(2)
    t1 = run(16)
    threads = threads_all
    return {"value": NI * nj * NT / t, "unit": "grid points/s", "cores": threads, "kind": "port",
            "serial_value": NI * 16 * NT / t1, "serial_sample": f"{NI}x16 rows x {NT} steps, 1 thread, {t1:.1f} s",
            "sample": f"{NI}x{nj} rows (centre of the 1440x720 grid) x {NT} steps, {t:.1f} s, "
                      f"oracle/aerobulk_oracle.c gcc -O2 -ffp-contract=off, OpenMP row blocks"}


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference algorithm's CPU implementation (the C oracle port: no Fortran
    compiler exists in this image) on all host cores; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    from aerobulk_b200 import synth
    from oracle.oracle import OracleSession

    threads = os.cpu_count() or 1
    budget = 150.0 / max(1, args.steps + args.warmup)      # seconds per step
    probe_nj = 8
    fp = synth.fields(NI, NJ, j0=356, j1=356 + probe_nj)

    def session(f, rsw):
        s = OracleSession(threads=threads)
        for jt in range(1, NT + 1):
            s.model(jt, NT, ALGO, ZT, ZU, *[f[k] for k in IN_KEYS], Niter=NB_ITER, l_use_skin=True,
                    rad_sw=rsw[jt - 1], rad_lw=f["rad_lw"])

    rswp = [synth.rad_sw_hour(NI, NJ, jt, j0=356, j1=356 + probe_nj) for jt in range(1, NT + 1)]
    session(fp, rswp)
    t0 = time.perf_counter()
    session(fp, rswp)
    rate = NI * probe_nj * NT / (time.perf_counter() - t0)
    nj = int(min(NJ, max(probe_nj, round(rate * budget / (NI * NT)))))
    j0 = (NJ - nj) // 2
    f = synth.fields(NI, NJ, j0=j0, j1=j0 + nj)
    rsw = [synth.rad_sw_hour(NI, NJ, jt, j0=j0, j1=j0 + nj) for jt in range(1, NT + 1)]
    for _ in range(args.warmup):
        session(f, rsw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        session(f, rsw)
    dt = time.perf_counter() - t0
    value = NI * nj * NT * args.steps / dt
    sample = (f"each step = {NI}x{nj} rows (centre of the 1440x720 grid) x {NT} hourly calls; "
              f"oracle/aerobulk_oracle.c (C port of the reference Fortran, gcc -O2 -ffp-contract=off), {threads} OpenMP threads")
    line = {"impl": "reference", "metric": "grid points/s (FP64)", "value": value, "unit": "grid points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "grid points/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "grid points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    import aerobulk_b200 as ab
    from aerobulk_b200 import model as abm
    from aerobulk_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the aerobulk_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    abm.set_device(local_rank)
    ab.reset()
    # a dedicated (non-default) torch stream: the library launches on it, torch.cuda.Event times it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ab.set_stream(stream.cuda_stream)

    n = NI * NJ
    njg = NJ * world
    j0 = rank * NJ
    f = synth.fields(NI, njg, j0=j0, j1=j0 + NJ)
    rsw_h = [synth.rad_sw_hour(NI, njg, jt, j0=j0, j1=j0 + NJ) for jt in range(1, NT + 1)]

    def pinned(a):
        t = torch.empty(n, dtype=torch.float64).pin_memory()
        t.numpy()[:] = np.ravel(a, order="F")
        return t

    # the model keeps its step-invariant fields in ONE pinned slab (sst .. slp, rad_lw) and its outputs in another:
    # the library then moves each pipeline chunk with one pitched copy per slab ("faster paths")
    in_slab = torch.empty((len(IN_KEYS) + 1, n), dtype=torch.float64).pin_memory()
    host = {}
    for i, k in enumerate(IN_KEYS + ("rad_lw",)):
        in_slab[i].numpy()[:] = np.ravel(f[k], order="F")
        host[k] = in_slab[i]
    host_rsw = [pinned(a) for a in rsw_h]
    # two device copies of the step-invariant inputs, used alternately, so that consecutive calls never
    # re-read the same lines (per call: 8 inputs + 6 outputs + 4 state arrays R/W = 150 MB > 126 MB L2)
    devin = [{k: host[k].to(dev, non_blocking=True) for k in IN_KEYS + ("rad_lw",)} for _ in range(2)]
    dev_rsw = [t.to(dev, non_blocking=True) for t in host_rsw]
    out = {k: torch.empty(n, dtype=torch.float64, device=dev) for k in OUT_KEYS}
    torch.cuda.synchronize()

    ops = abm.stats_reduce_ops()
    op_t = torch.from_numpy(ops).to(dev)

    def global_init(src):
        st = torch.from_numpy(abm.init_local_stats(*[src[k] for k in IN_KEYS], rad_lw=src["rad_lw"])).to(dev)
        big = torch.full_like(st, float("inf"))
        s_sum = torch.where(op_t == 0, st, torch.zeros_like(st))
        s_min = torch.where(op_t == 1, st, big)
        s_max = torch.where(op_t == 2, st, -big)
        dist.all_reduce(s_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(s_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(s_max, op=dist.ReduceOp.MAX)
        g = torch.where(op_t == 0, s_sum, torch.where(op_t == 1, s_min, s_max))
        abm.init_from_stats(NT, ALGO, True, True, g.cpu().numpy())

    ev_pairs = []

    def session_device(timed: bool):
        for jt in range(1, NT + 1):
            src = devin[jt & 1]
            if jt == 1 and world > 1:
                global_init(src)
            if timed and jt > 1:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            ab.aerobulk_model_device(jt, NT, ALGO, ZT, ZU, *[src[k] for k in IN_KEYS], out=out, Niter=NB_ITER,
                                     l_use_skin=True, rad_sw=dev_rsw[jt - 1], rad_lw=src["rad_lw"], shape=(NI, NJ))
            if timed and jt > 1:
                e1.record(stream)
                ev_pairs.append((e0, e1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(3, args.warmup)):
        session_device(False)
    barrier()
    abm.reset_launch_count()
    t_wall0 = time.perf_counter()
    E0, E1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    E0.record(stream)
    for _ in range(args.steps):
        session_device(True)
    E1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    launches = ab.launch_count()
    ms = torch.tensor([E0.elapsed_time(E1)], dtype=torch.float64, device=dev)
    kern_ms = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in ev_pairs]))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    ms, kern_ms = float(ms.item()), float(kern_ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = world * n * NT * args.steps / (ms * 1e-3)

    # ---------------- end to end through the host-array API (`e2e`): pinned host inputs, H2D + D2H inside
    out_slab = torch.empty((len(OUT_KEYS), n), dtype=torch.float64).pin_memory()
    host_out = {k: out_slab[i] for i, k in enumerate(OUT_KEYS)}
    np_in = {k: v.numpy().reshape((NI, NJ), order="F") for k, v in host.items()}
    np_rsw = [t.numpy().reshape((NI, NJ), order="F") for t in host_rsw]
    np_out = {k: v.numpy().reshape((NI, NJ), order="F") for k, v in host_out.items()}

    def session_host():
        for jt in range(1, NT + 1):
            if jt == 1 and world > 1:
                global_init(devin[1])
            ab.aerobulk_model(jt, NT, ALGO, ZT, ZU, *[np_in[k] for k in IN_KEYS], Niter=NB_ITER, l_use_skin=True,
                              rad_sw=np_rsw[jt - 1], rad_lw=np_in["rad_lw"], out=np_out)

    for _ in range(2):
        session_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        session_host()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * n * NT * args.steps / float(e2e_s.item())

    # ---------------- the same call with PAGEABLE arrays (a caller that only relinked): reported beside `e2e`, N=1 only
    e2e_pageable = None
    if world == 1:
        pg_in = {k: np.array(v, order="F") for k, v in np_in.items()}
        pg_rsw = [np.array(v, order="F") for v in np_rsw]
        pg_out = {k: np.zeros((NI, NJ), order="F") for k in OUT_KEYS}

        def session_pageable():
            for jt in range(1, NT + 1):
                ab.aerobulk_model(jt, NT, ALGO, ZT, ZU, *[pg_in[k] for k in IN_KEYS], Niter=NB_ITER, l_use_skin=True,
                                  rad_sw=pg_rsw[jt - 1], rad_lw=pg_in["rad_lw"], out=pg_out)

        session_pageable()
        torch.cuda.synchronize()
        reps = min(args.steps, 5)
        t0 = time.perf_counter()
        for _ in range(reps):
            session_pageable()
        torch.cuda.synchronize()
        e2e_pageable = n * NT * reps / (time.perf_counter() - t0)
        del pg_in, pg_rsw, pg_out

    # ---------------- optional global diagnostics of the last device-resident step (outside every timed region):
    # row-block sums / minima / maxima on the device, combined across ranks by NCCL all-reduces of 19 doubles
    dvec = torch.from_numpy(abm.flux_diagnostics(out)).to(dev)
    if world > 1:
        dop = torch.from_numpy(abm.diag_reduce_ops()).to(dev)
        big = torch.full_like(dvec, float("inf"))
        d_sum = torch.where(dop == 0, dvec, torch.zeros_like(dvec))
        d_min = torch.where(dop == 1, dvec, big)
        d_max = torch.where(dop == 2, dvec, -big)
        dist.all_reduce(d_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(d_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(d_max, op=dist.ReduceOp.MAX)
        dvec = torch.where(dop == 0, d_sum, torch.where(dop == 1, d_min, d_max))
    diagnostics = abm.diagnostics_summary(dvec.cpu().numpy())

    # ---------------- rooflines
    hbm_peak, hbm_src = measured_peaks()
    bytes_pt = ab.bytes_per_point(ALGO, SKIN)
    work_pt = ab.work_per_point(ALGO, SKIN, NB_ITER)
    fp64_peak = ab.measure_fp64_peak() if rank == 0 else 0.0
    traffic, fp64_inst = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["flux_kernel<COARE3P6,skin,zt!=zu>@1440x720"]
        traffic = tj["traffic"]
        fp64_inst = tj.get("fp64_thread_inst_per_launch")
    except Exception:
        pass
    achieved_gbs = bytes_pt * n / (kern_ms * 1e-3) / 1e9
    achieved_fp64 = work_pt * n / (kern_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": "grid points/s (FP64)", "value": value, "unit": "grid points/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "algo": ALGO, "skin": SKIN, "grid_per_gpu": [NI, NJ], "nt": NT,
                   "nb_iter": NB_ITER, "points_per_step_per_gpu": n * NT,
                   "l2": "per call 8 inputs + 6 outputs + 4 state arrays R/W = 150 MB > 126 MB L2; "
                         "step-invariant inputs alternate between two device copies, 24 distinct rad_sw fields",
                   "sharding": "latitude row blocks, one rank per GPU, no data-path collective"},
        "e2e": {"value": e2e_value, "unit": "grid points/s", "h2d_bytes_per_step": 8 * 8 * n * NT,
                "d2h_bytes_per_step": 6 * 8 * n * NT,
                "how": "aerobulk_gpu_model (host-array C ABI) with pinned host buffers: the kernel loads its inputs from and stores its outputs to them directly over PCIe (zero-copy, every byte crosses once per call); jt=1 of each session goes through the staged chunked H2D|kernel|D2H pipeline"},
        "e2e_pageable": {"value": e2e_pageable, "unit": "grid points/s",
                         "how": "same calls with ordinary (pageable) numpy arrays: the library moves them through its own "
                                "pinned slab on host copy threads around a zero-copy kernel (DESIGN.md 4); N=1 only"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "diagnostics": diagnostics,
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": traffic,
                     "kernel": "flux_kernel<COARE3P6,skin,zt!=zu>", "avg_launch_ms": kern_ms,
                     "algorithmic_bytes_per_point": bytes_pt, "peak_source": hbm_src,
                     "note": "this path is FP64-pipe bound, not HBM bound (SURVEY 8d): see roofline_fp64"},
        "roofline_fp64": {"bound": "fp64", "achieved": achieved_fp64 / 1e12, "peak": fp64_peak / 1e12,
                          "unit": "T FP64-pipe instr/s", "frac": (achieved_fp64 / fp64_peak) if fp64_peak > 0 else None,
                          "work_per_point": work_pt,
                          "executed_frac": (fp64_inst / (kern_ms * 1e-3) / fp64_peak) if (fp64_inst and fp64_peak > 0) else None,
                          "note": "achieved = algorithmic FP64-pipe instruction equivalents of the REFERENCE arithmetic "
                                  "(SURVEY 8d: fx + nb_iter*it) x points / kernel time; peak = DFMA chain measured live; "
                                  "frac > 1 because the kernel needs ~2.7x fewer FP64 instructions than that arithmetic "
                                  "(own exp/log/atan, no pow, fast division: DESIGN.md 3.1); executed_frac = FP64-pipe "
                                  "instructions actually executed per launch (ncu, profiles/traffic.json) / time / peak"},
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(os.cpu_count() or 1)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
