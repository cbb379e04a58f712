"""Throughput of the station time-series launch (aerobulk_gpu_series, host arrays) versus the number of stations,
with the oracle's serial and threaded figures beside it.  COARE 3.6 + skin, nb_iter = 20 as the reference program."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aerobulk_b200 as ab
from aerobulk_b200 import synth

Nt = int(os.environ.get("SERIES_NT", "240"))
algo = os.environ.get("SERIES_ALGO", "coare3p6")
want = ("QL", "QH", "TAU", "Ts", "dT_wl", "Hz_wl")
ab.set_nb_iter(20)
print(f"algo {algo} skin nb_iter 20 Nt {Nt}")
print(f"{'stations':>9s} {'ms':>10s} {'station-records/s':>18s}")
for S in (1, 32, 1024, 148 * 64, 148 * 64 * 8, 148 * 64 * 32):
    d = synth.station_series(Nt, S)
    ab.series(algo, 2.0, 10.0, **d, want=want)
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        ab.series(algo, 2.0, 10.0, **d, want=want)
        best = min(best, time.perf_counter() - t)
    print(f"{S:9d} {best * 1e3:10.2f} {Nt * S / best:18.4g}")
# pinned host arrays: zero-copy (the kernel reads / writes the caller's memory over PCIe)
import ctypes as C
import torch
print("pinned host arrays (zero-copy)")
for S in (1024, 148 * 64 * 8, 148 * 64 * 32):
    d = synth.station_series(Nt, S)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
    keep = {k: pin(v) for k, v in d.items() if k != "isecday_utc"}
    outs = {k: torch.zeros((Nt, S), dtype=torch.float64).pin_memory() for k in want}
    arr = (C.c_void_p * len(ab.SERIES_OUT))(*[outs[k].data_ptr() if k in outs else None for k in ab.SERIES_OUT])
    isd = np.ascontiguousarray(d["isecday_utc"], dtype=np.int32)
    call = lambda: ab.lib().aerobulk_gpu_series(algo.encode(), Nt, S, 2.0, 10.0, isd.ctypes.data, keep["lon"].data_ptr(),
        keep["sst"].data_ptr(), keep["t_zt"].data_ptr(), keep["hum_zt"].data_ptr(), 0, keep["wind"].data_ptr(),
        keep["slp"].data_ptr(), keep["rad_sw"].data_ptr(), keep["rad_lw"].data_ptr(), 1, C.cast(arr, C.c_void_p), 0)
    assert call() == 0
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        call()
        best = min(best, time.perf_counter() - t)
    print(f"{S:9d} {best * 1e3:10.2f} {Nt * S / best:18.4g}")
# device-resident: the kernel alone (isecday_utc upload and the stress flag read-back included)
print("device-resident (aerobulk_gpu_series on_device=1)")
for S in (1, 1024, 148 * 64, 148 * 64 * 8, 148 * 64 * 32):
    d = synth.station_series(Nt, S)
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items() if k != "isecday_utc"}
    out = {k: torch.empty((Nt, S), dtype=torch.float64, device="cuda") for k in want}
    args = (algo, 2.0, 10.0, d["isecday_utc"], dev["lon"], dev["sst"], dev["t_zt"], dev["hum_zt"], dev["wind"], dev["slp"],
            dev["rad_sw"], dev["rad_lw"], out)
    ab.series_device(*args)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        ab.series_device(*args)
        best = min(best, time.perf_counter() - t)
    print(f"{S:9d} {best * 1e3:10.2f} {Nt * S / best:18.4g}")
if os.environ.get("SERIES_CPU", "1") == "1":
    from oracle.oracle import OracleSession
    for S, thr in ((1, 1), (256, 1), (4096, os.cpu_count())):
        d = synth.station_series(Nt, S)
        o = OracleSession(threads=thr)
        o.set_nb_iter(20)
        t = time.perf_counter()
        o.series(algo, 2.0, 10.0, **d)
        dt = time.perf_counter() - t
        print(f"oracle S={S} threads={thr}: {dt * 1e3:.1f} ms, {Nt * S / dt:.4g} station-records/s")
