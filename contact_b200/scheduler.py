"""Batched case scheduler: static sharding of independent contact cases over ranks + the final gather.

SURVEY.md 8(e): the unit is one contact case; cases shard with no exchange inside a solve; the only collective is one
gather of per-case results at the end.  Used by bench.py (NCCL) and by the world_size-2 gloo tests (CPU tensors).
"""
import torch
import torch.distributed as dist


def partition(ncase, world):
    """Contiguous block partition of case indices: list of (lo, hi) per rank, sizes differ by at most one."""
    base, rem = divmod(ncase, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def my_range(ncase, rank, world):
    return partition(ncase, world)[rank]


def gather_case_results(local, ncase, group=None):
    """All-gather per-case result rows (local: (n_local, k) tensor) into the (ncase, k) table in case order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    parts = partition(ncase, world)
    nmax = max(hi - lo for lo, hi in parts)
    pad = torch.zeros((nmax, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:hi - lo] for r, (lo, hi) in enumerate(parts)], dim=0)
