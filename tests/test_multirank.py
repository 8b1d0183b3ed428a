"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chain partitioning and the end-of-run collectives."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from fk_mc_b200 import parallel, stats


def test_partition_chains():
    for total, world in [(4096, 8), (10, 3), (7, 8), (0, 2), (1024, 1)]:
        spans = [parallel.partition_chains(total, world, r) for r in range(world)]
        assert sum(n for _, n in spans) == total
        pos = 0
        for c0, n in spans:
            assert c0 == pos
            pos += n
        assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    with pytest.raises(ValueError):
        parallel.partition_chains(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, n_meas, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c0, n = parallel.partition_chains(total, world, rank)
    # synthetic per-chain series: value encodes (measurement, global chain id) so ordering errors are visible
    local = np.array([[1000.0 * m + (c0 + c) for c in range(n)] for m in range(n_meas)]).reshape(n_meas, n)
    full = parallel.gather_series(local, total).numpy()
    expect = np.array([[1000.0 * m + c for c in range(total)] for m in range(n_meas)])
    ok = np.array_equal(full, expect)
    # moment reduction == moments of the concatenated per-rank series, level by level
    mom = parallel.reduce_moments(local, levels=4)
    ref = np.zeros((4, 3))
    for r in range(world):
        rc0, rn = parallel.partition_chains(total, world, r)
        x = np.array([[1000.0 * m + (rc0 + c) for c in range(rn)] for m in range(n_meas)]).reshape(n_meas, rn).T.reshape(-1)  # chain-major
        for lv in range(4):
            if x.size:
                ref[lv] += (x.size, x.sum(), (x * x).sum())
            k = x.size // 2
            x = 0.5 * (x[0:2 * k:2] + x[1:2 * k:2])
    ok = ok and np.allclose(mom, ref)
    # binned statistics of the gathered series are identical on every rank
    rows = stats.accumulate_binning(full.reshape(-1), 3)
    t = torch.tensor([r[1] for r in rows])
    ts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(ts, t)
    ok = ok and all(torch.equal(ts[0], x) for x in ts)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_series_world2(tmp_path, total):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, 5, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "ok%d" % r)).read() == "1"
