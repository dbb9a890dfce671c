"""Multi-GPU plumbing: problems are independent, so the batch index range [0, B) is cut into
contiguous chunks, one per rank (one process per GPU), with NO collective on the solve path;
the only exchange is the final gather of the small per-problem results."""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous chunk [lo, hi) of rank ``rank``; chunks differ by at most one problem."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local, B, rank, world, dist=None, device=None):
    """All-gather a dict of per-problem arrays (first axis = local problems) into full-batch
    arrays ordered by global problem index.  ``dist`` is torch.distributed (initialised)."""
    if world == 1 or dist is None:
        return {k: np.asarray(v) for k, v in local.items()}
    import torch
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    out = {}
    for k, v in local.items():
        v = np.ascontiguousarray(v)
        pad = max(sizes)
        buf = torch.zeros((pad,) + v.shape[1:], dtype=torch.from_numpy(v).dtype, device=device)
        buf[:v.shape[0]] = torch.from_numpy(v).to(device)
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        out[k] = np.concatenate([p[:n].cpu().numpy() for p, n in zip(parts, sizes)], axis=0)
    return out
