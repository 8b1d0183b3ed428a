"""End-of-run collective through the C ABI (fkmc_comm_init / fkmc_gather_series): two ranks, one GPU each, all-gather of the
energy series from the chain engine's device buffers over NCCL.  Needs two GPUs (gpurun --gpus 2); skipped otherwise.
Replaces the root-0 gathers of src/measures/energy.cpp:32-47."""
import os

import numpy as np
import pytest

import fk_mc_b200 as fk

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import ctypes
        n = ctypes.c_int(0)
        return n.value if ctypes.CDLL("libcudart.so.12").cudaGetDeviceCount(ctypes.byref(n)) else n.value
    except OSError:
        return 0


def _rank(rank, world, tmp, nch, nsw):
    idf = os.path.join(tmp, "nccl_id")
    if rank == 0:
        with open(idf + ".tmp", "wb") as fh:
            fh.write(fk.nccl_unique_id())
        os.rename(idf + ".tmp", idf)
    else:
        import time
        while not os.path.exists(idf):
            time.sleep(0.05)
    uid = open(idf, "rb").read()
    c = fk.Context("cubic2d", 8, max_batch=nch, device=rank)
    c.chain_init(nch, 4.0, 4.0, seed=32167, chain0=rank * nch, sweep_len=16, ntherm_sweeps=1, max_sweeps=nsw + 1)
    c.chain_run_sweeps(nsw + 1)
    c.comm_init(uid, world, rank)
    g = c.gather_series()
    np.save(os.path.join(tmp, "gathered%d.npy" % rank), np.stack([g["energies"], g["d2energies"], g["c_energies"]]))
    c.close()


def test_single_rank_gather_is_the_local_series():
    c = fk.Context("cubic2d", 8, max_batch=4)
    c.chain_init(4, 4.0, 4.0, seed=1, max_sweeps=3)
    c.chain_run_sweeps(3)
    g, s = c.gather_series(), c.chain_get_series()
    assert g["n_measured"] == 2 and np.array_equal(g["energies"], s["energies"]) and np.array_equal(g["c_energies"], s["c_energies"])
    c.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_rank_device_gather(tmp_path):
    import multiprocessing as mp
    world, nch, nsw = 2, 6, 3
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_rank, args=(r, world, str(tmp_path), nch, nsw)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # the same 12 chains on one GPU: the gathered series must not depend on how the chains were split
    c = fk.Context("cubic2d", 8, max_batch=world * nch)
    c.chain_init(world * nch, 4.0, 4.0, seed=32167, sweep_len=16, ntherm_sweeps=1, max_sweeps=nsw + 1)
    c.chain_run_sweeps(nsw + 1)
    s = c.chain_get_series()
    ref = np.stack([s["energies"], s["d2energies"], s["c_energies"]])
    for r in range(world):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "gathered%d.npy" % r)), ref)
    c.close()
