"""Multi-GPU plumbing: controller instances are independent, so a batch is sharded contiguously over
the ranks (one process per GPU, one libbmpc handle per process) with no data-path collective.  The
only exchange is the optional all-gather of the computed moves Z̃ (or u) after a step, so that every
rank sees the whole batch's decisions (north_star); NCCL over NVLink on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous, balanced partition: first (n_total % world) ranks get one extra instance."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_moves(z_local, n_total=None):
    """All-gather the per-rank move arrays (n_local, n) into the full (n_total, n) array on every rank.
    ``z_local`` is a torch tensor (CUDA with the nccl backend, CPU with gloo).  Shards may differ by
    one row; they are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    n_local = z_local.shape[0]
    if n_total is None:
        t = torch.tensor([n_local], device=z_local.device)
        dist.all_reduce(t)
        n_total = int(t.item())
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((nmax,) + tuple(z_local.shape[1:]), dtype=z_local.dtype, device=z_local.device)
    pad[:n_local] = z_local
    out = torch.empty((world * nmax,) + tuple(z_local.shape[1:]), dtype=z_local.dtype, device=z_local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * nmax:r * nmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)
