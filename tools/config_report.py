"""Throughput of the five BASELINE.json configurations on one B200: device-resident (CUDA events), end to end through the
host-array API with pinned arrays, and the CPU oracle on the box's host cores (bounded row sample).  Point/s = grid
points per aerobulk_model call / time per call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aerobulk_b200 as ab
from aerobulk_b200 import synth
from oracle.oracle import OracleSession

IN = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
OUT = ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
NCPU = os.cpu_count() or 1


def fields(Ni, Nj, j0=0, j1=None, NjG=None):
    return synth.fields(Ni, NjG or Nj, j0=j0, j1=j1 if j1 is not None else Nj)


def device_time(algo, skin, f, nb_iter, Ni, Nj, steps=5):
    n = Ni * Nj
    dev = {k: torch.from_numpy(np.ravel(v, order="F").copy()).cuda() for k, v in f.items()}
    out = {k: torch.empty(n, dtype=torch.float64, device="cuda") for k in (OUT if skin else OUT[:5])}
    ab.reset(); ab.set_verbose(False); ab.set_stream(st.cuda_stream)
    kw = dict(Niter=nb_iter)
    if skin:
        kw.update(l_use_skin=True, rad_sw=dev["rad_sw"], rad_lw=dev["rad_lw"])
    ts = []
    for jt in range(1, steps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ab.aerobulk_model_device(jt, steps, algo, 2., 10., *[dev[k] for k in IN], out=out, shape=(Ni, Nj), **kw)
        e1.record(st)
        torch.cuda.synchronize()
        if jt > 1:
            ts.append(e0.elapsed_time(e1))
    ab.set_stream(None)
    return n / (np.median(ts) * 1e-3)


def e2e_time(algo, skin, f, nb_iter, Ni, Nj, steps=4):
    n = Ni * Nj
    pin = lambda a: torch.from_numpy(np.ravel(a, order="F").copy()).pin_memory()
    h = {k: pin(v) for k, v in f.items()}
    v = lambda t: t.numpy().reshape((Ni, Nj), order="F")
    out = {k: v(torch.empty(n, dtype=torch.float64).pin_memory()) for k in (OUT if skin else OUT[:5])}
    ab.reset(); ab.set_verbose(False)
    kw = dict(Niter=nb_iter)
    if skin:
        kw.update(l_use_skin=True, rad_sw=v(h["rad_sw"]), rad_lw=v(h["rad_lw"]))
    ts = []
    for jt in range(1, steps + 1):
        t0 = time.perf_counter()
        ab.aerobulk_model(jt, steps, algo, 2., 10., *[v(h[k]) for k in IN], out=out, **kw)
        if jt > 1:
            ts.append(time.perf_counter() - t0)
    return n / np.median(ts)


def cpu_time(algo, skin, f, nb_iter, Ni, Nj, threads, budget_s=6.0):
    rows = max(1, min(Nj, int(budget_s * (2.5e6 if threads > 1 else 1.8e5) * (0.3 if skin else 1.0) * 5 / nb_iter / Ni)))
    j0 = (Nj - rows) // 2
    g = {k: np.asfortranarray(v[:, j0:j0 + rows]) for k, v in f.items()}
    o = OracleSession(threads=threads)
    kw = dict(Niter=nb_iter)
    if skin:
        kw.update(l_use_skin=True, rad_sw=g["rad_sw"], rad_lw=g["rad_lw"])
    o.model(1, 3, algo, 2., 10., *[g[k] for k in IN], **kw)
    t0 = time.perf_counter()
    o.model(2, 3, algo, 2., 10., *[g[k] for k in IN], **kw)
    return Ni * rows / (time.perf_counter() - t0), rows


def line(tag, algo, skin, Ni, Nj, nb_iter=5, f=None, cpu=True, e2e=True):
    f = f if f is not None else fields(Ni, Nj)
    d = device_time(algo, skin, f, nb_iter, Ni, Nj)
    e = e2e_time(algo, skin, f, nb_iter, Ni, Nj) if e2e else float("nan")
    if cpu:
        c1, r1 = cpu_time(algo, skin, f, nb_iter, Ni, Nj, 1)
        ca, ra = cpu_time(algo, skin, f, nb_iter, Ni, Nj, NCPU)
        cs = f"{c1 / 1e6:7.3f} (1 thread, {r1} rows)  {ca / 1e6:7.2f} ({NCPU} threads, {ra} rows)"
    else:
        cs = "-"
    print(f"{tag:4s} {algo:9s} skin={int(skin)} nb_iter={nb_iter:2d} {Ni:6d}x{Nj:<5d} device {d / 1e9:7.3f} Gpt/s   e2e {e / 1e9:6.3f} Gpt/s   CPU oracle Mpt/s: {cs}", flush=True)


print(f"one B200, host cores {NCPU}; device = aerobulk_gpu_model_device (CUDA events, steps 2..), e2e = aerobulk_gpu_model with pinned arrays (zero-copy), wall clock")
line("C1", "ncar", False, 360, 180)
line("C2", "coare3p6", True, 1440, 720)
line("C3", "ecmwf", True, 4320, 2160)
for algo in ("andreas", "coare3p0"):
    for nb in (5, 10, 20, 30):
        line("C4", algo, False, 4320, 2160, nb_iter=nb, cpu=(nb in (5, 30)), e2e=(nb == 5))
f5 = fields(12960, 810, j0=2835, j1=3645, NjG=6480)      # the equatorial shard of eight of the 1/36 deg grid
for algo in ("ncar", "andreas", "coare3p0", "coare3p6", "ecmwf"):
    line("C5/8", algo, False, 12960, 810, f=f5, cpu=(algo == "ncar"))
for algo in ("coare3p6", "ecmwf"):
    line("C5/8", algo, True, 12960, 810, f=f5, cpu=False)
