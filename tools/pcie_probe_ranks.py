"""Concurrent host<->device bandwidth of N GPUs of one box: what limits the end-to-end (host-array) path as GPUs are added.

    python tools/pcie_probe_ranks.py                                  one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_ranks.py

Every rank copies 256 MiB pinned buffers to / from its own GPU, all ranks at the same time (barrier before each leg);
legs: H2D alone, D2H alone, both directions at once.  Rank 0 prints one line per leg: GB/s per GPU (min / mean / max over
the ranks) and the aggregate over the box, plus the topology facts that matter (NUMA nodes, CPUs, affinity).
A host memory copy test (numpy, one thread per rank) gives the host-DRAM figure the pageable path shares."""
import os, subprocess, time
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 32 * 1024 * 1024            # 256 MiB of float64
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
h_in.fill_(1.0)
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.ones(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def leg(fn, reps=8):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    return dt


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world == 1:
        return [x]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


if rank == 0:
    try:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        numa = subprocess.run(["bash", "-c", "lscpu | grep -E 'NUMA node|^CPU\\(s\\)|Model name|Socket'"], capture_output=True, text=True).stdout
        print(f"# {world} GPU(s); affinity of rank 0: {sorted(os.sched_getaffinity(0))[:4]}.. ({len(os.sched_getaffinity(0))} CPUs)")
        print("# " + numa.strip().replace("\n", "\n# "))
        print("# " + "\n# ".join(l for l in topo.splitlines()[:world + 1]))
    except Exception as exc:
        print("# topology unavailable:", exc)
for name, fn, dirs in (("H2D alone", h2d, 1), ("D2H alone", d2h, 1), ("H2D + D2H together", both, 2)):
    dt = leg(fn)
    per = gather(n * 8 * dirs / dt / 1e9)
    if rank == 0:
        print(f"{name:22s}: per GPU min {min(per):6.1f} mean {np.mean(per):6.1f} max {max(per):6.1f} GB/s"
              f"{' (sum of both directions)' if dirs == 2 else ''}; box aggregate {sum(per):7.1f} GB/s")
# host DRAM: one streaming copy per rank at the same time (what the pageable bounce path shares with the DMA engines)
a = np.ones(n)
b = np.empty(n)
barrier()
t0 = time.perf_counter()
for _ in range(4):
    np.copyto(b, a)
dt = (time.perf_counter() - t0) / 4
per = gather(2 * n * 8 / dt / 1e9)
if rank == 0:
    print(f"{'host memcpy (1 thread)':22s}: per rank min {min(per):6.1f} mean {np.mean(per):6.1f} GB/s (read + write); aggregate {sum(per):7.1f} GB/s")
if world > 1:
    dist.destroy_process_group()
