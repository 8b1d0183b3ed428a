"""Multi-GPU layout of the Markov chains: one process per GPU, chains partitioned contiguously,
no traffic while sampling, one collective at the end.

The reference runs one chain per MPI rank with seed SEED + rank (src/mc_metropolis.cpp:25) and would
gather the per-rank series on rank 0 (src/measures/energy.cpp:32-47).  Here chain c of the job is
that "rank c"; GPU g owns chains [g*C/G, (g+1)*C/G), so results do not depend on the GPU count.
torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def partition_chains(total_chains, world_size, rank):
    """Contiguous split; the first (total % world) ranks get one extra chain.  Returns (chain0, n_local)."""
    if total_chains < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad partition arguments")
    base, extra = divmod(total_chains, world_size)
    n_local = base + (1 if rank < extra else 0)
    chain0 = rank * base + min(rank, extra)
    return chain0, n_local


def gather_series(local, total_chains, group=None, device=None):
    """All-gather per-chain series [n_measured, n_local] into [n_measured, total_chains] (chain order = global id).

    `local` is a numpy array or a torch tensor (CPU for gloo, CUDA for NCCL).  Mirrors the reference's
    gather of `_energies` (src/measures/energy.cpp:36-38), with every rank receiving the result."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(local)
    if device is not None:
        t = t.to(device)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    n_meas = t.shape[0]
    counts = [partition_chains(total_chains, world, r)[1] for r in range(world)]
    width = max(counts)
    padded = torch.zeros((n_meas, width), dtype=t.dtype, device=t.device)
    padded[:, : t.shape[1]] = t
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded.contiguous(), group=group)
    return torch.cat([o[:, :c] for o, c in zip(out, counts)], dim=1)


def reduce_moments(local, levels=16, group=None, device=None):
    """Sum over ranks of per-bin-level accumulators [levels, 3] = (n, sum x, sum x^2) of a pooled series
    (`local`: one chain's 1-D series or [measurement][chain]; a bin never mixes chains only when the per-chain length is a
    multiple of the bin width -- the same caveat the reference's concatenated ranks have).

    The cheap alternative to gather_series when only /stats is wanted: a few KB per observable."""
    import torch
    import torch.distributed as dist

    from .stats import pool_chains
    x = pool_chains(local)  # [measurement][chain] -> chain-major: bins run along Monte Carlo time inside a chain
    acc = np.zeros((levels, 3))
    cur = x
    for lv in range(levels):
        if cur.size == 0:
            break
        acc[lv] = (cur.size, cur.sum(), (cur * cur).sum())
        m = cur.size // 2
        cur = 0.5 * (cur[0:2 * m:2] + cur[1:2 * m:2])
    t = torch.from_numpy(acc)
    if device is not None:
        t = t.to(device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


class DeviceArray:
    """Zero-copy view of a device buffer owned by libfkmc_b200, consumable by torch.as_tensor(..., device='cuda')."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 3,
                                         "strides": None}
