"""N>1 host logic on CPU (gloo, world_size 2): the row-block sharded AEROBULK_INIT.

Each rank owns a latitude row block.  It computes the statistics vector of its block (on the GPU this
is aerobulk_gpu_init_local_stats; here a numpy statement of the same layout), the ranks combine them
with the per-entry reduce op the library publishes (sum / min / max all-reduce), and every rank runs
AEROBULK_INIT's decisions through aerobulk_gpu_init_from_stats (pure host code of the C ABI).
The decisions must equal those of the single-process oracle on the whole grid."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aerobulk_b200 import synth

NST = 64


def numpy_stats(f, with_rad):
    """Same layout as include/aerobulk_gpu.h: [n_unmasked, n, then per field msum mmin mmax min max]."""
    wnd = np.sqrt(f["U_zu"] * f["U_zu"] + f["V_zu"] * f["V_zu"])
    m = (f["sst"] >= 270) & (f["sst"] <= 320) & (f["t_zt"] >= 180) & (f["t_zt"] <= 330) & \
        (f["slp"] >= 80000) & (f["slp"] <= 110000) & (wnd <= 50)
    if with_rad:
        m &= (f["rad_lw"] >= 0) & (f["rad_lw"] <= 750)
    st = np.zeros(NST)
    st[0], st[1] = m.sum(), m.size
    rl = f["rad_lw"] if with_rad else np.zeros_like(wnd)
    fields = [f["sst"], f["t_zt"], f["slp"], f["U_zu"], f["V_zu"], wnd, f["hum_zt"], rl, rl]
    for k, v in enumerate(fields):
        b = 2 + 5 * k
        st[b] = v[m].sum() if m.any() else 0.0
        st[b + 1] = v[m].min() if m.any() else np.finfo(float).max
        st[b + 2] = v[m].max() if m.any() else -np.finfo(float).max
        st[b + 3], st[b + 4] = v.min(), v.max()
    return st


def _worker(rank, world, port, hum, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aerobulk_b200 import model as abm
        abm.lib()
        abm.reset()
        Ni, Nj = 64, 48
        j0, j1 = rank * Nj // world, (rank + 1) * Nj // world
        f = synth.fields(Ni, Nj, j0=j0, j1=j1, humidity=hum)
        st = torch.from_numpy(numpy_stats(f, True))
        ops = torch.from_numpy(abm.stats_reduce_ops())
        big = torch.full_like(st, float("inf"))
        s_sum = torch.where(ops == 0, st, torch.zeros_like(st))
        s_min = torch.where(ops == 1, st, big)
        s_max = torch.where(ops == 2, st, -big)
        dist.all_reduce(s_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(s_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(s_max, op=dist.ReduceOp.MAX)
        g = torch.where(ops == 0, s_sum, torch.where(ops == 1, s_min, s_max)).numpy()
        abm.init_from_stats(24, "coare3p6", True, True, g)
        ret[rank] = (g.copy(), abm.humidity_type(), abm.use_skin())
        # a unit error on ONE rank's block must stop EVERY rank (global statistics)
        abm.reset()
        f2 = dict(f)
        if rank == 1:
            f2["slp"] = f["slp"] / 100.0     # hPa on rank 1 only -> those points are masked, mean still fine
            f2["sst"] = f["sst"] - 273.15    # and deg C -> masked too
        st = torch.from_numpy(numpy_stats(f2, True))
        s_sum = torch.where(ops == 0, st, torch.zeros_like(st))
        s_min = torch.where(ops == 1, st, big)
        s_max = torch.where(ops == 2, st, -big)
        dist.all_reduce(s_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(s_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(s_max, op=dist.ReduceOp.MAX)
        g2 = torch.where(ops == 0, s_sum, torch.where(ops == 1, s_min, s_max)).numpy()
        abm.init_from_stats(24, "ncar", None, False, g2)     # rank 1's block fully masked: still fine globally
        assert g2[0] == (Ni * Nj) // 2 and g2[1] == Ni * Nj
        ret[rank + world] = "masked-block-ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("hum", ["sh", "rh"])
def test_sharded_init_two_ranks(hum):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, hum, ret), nprocs=2, join=True)
    full = synth.fields(64, 48, humidity=hum)
    ref = numpy_stats(full, True)
    for rank in (0, 1):
        g, h, skin = ret[rank]
        assert h == hum and skin is True
        ops = None
        assert g[0] == ref[0] and g[1] == ref[1]
        for k in range(9):
            b = 2 + 5 * k
            assert g[b] == pytest.approx(ref[b], rel=1e-13)          # sums: order of addition differs
            assert np.array_equal(g[b + 1:b + 5], ref[b + 1:b + 5])  # min / max: exact
        assert ret[rank + 2] == "masked-block-ok"
    # same decisions as the single-process oracle on the whole grid
    from oracle.oracle import OracleSession
    o = OracleSession()
    o.model(1, 24, "coare3p6", 2.0, 10.0, *[full[k] for k in ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")],
            l_use_skin=True, rad_sw=full["rad_sw"], rad_lw=full["rad_lw"])
    assert o.humidity_type == hum and o.use_skin


def test_row_block_partition_of_the_oracle_is_exact():
    """Points are independent: the oracle on row blocks equals the oracle on the full grid (the
    property the GPU partition-invariance test checks bit for bit on the device)."""
    from oracle.oracle import OracleSession
    Ni, Nj = 40, 24
    keys = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
    f = synth.fields(Ni, Nj)
    full = OracleSession().model(1, 1, "ecmwf", 2.0, 10.0, *[f[k] for k in keys])
    for j0, j1 in ((0, 7), (7, 24)):
        b = synth.fields(Ni, Nj, j0=j0, j1=j1)
        blk = OracleSession().model(1, 1, "ecmwf", 2.0, 10.0, *[b[k] for k in keys])
        for k in full:
            assert np.array_equal(blk[k], full[k][:, j0:j1])
