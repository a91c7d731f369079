"""Row-sharding of the candidate pool over the GPUs of one node.

Every candidate row depends only on its own tree and on the (replicated) evaluated-set state, so
rank r takes the contiguous block ``shard_bounds(N, P, r)`` and the only exchange is ONE all-gather
of the per-rank acquisition slice (8 bytes per candidate).  ``torch.distributed`` provides the
plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests of this host-side logic).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row block of rank ``rank``: sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def padded_shard(n: int, world: int) -> int:
    """Slots per rank in the gathered buffer (all-gather needs equal contributions)."""
    return (n + world - 1) // world


class ShardedAcquisition:
    """``values = f(trees)`` evaluated shard-wise and gathered on every rank.

    ``compute(trees_local) -> float64 tensor [n_local]`` is the per-rank work (on the GPU box:
    ``AcquisitionEngine.acquire_device(...)[2]``).  ``trees`` is the FULL pool on every rank (host
    or device); each rank slices its own block, so no scatter is needed.
    """

    def __init__(self, compute: Callable[[torch.Tensor], torch.Tensor], group: Optional[dist.ProcessGroup] = None):
        self.compute = compute
        self.group = group

    def __call__(self, trees: torch.Tensor) -> torch.Tensor:
        if not (dist.is_available() and dist.is_initialized()):
            return self.compute(trees)
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        n = int(trees.shape[0])
        lo, hi = shard_bounds(n, world, rank)
        local = self.compute(trees[lo:hi])
        slot = padded_shard(n, world)
        send = torch.zeros(slot, dtype=local.dtype, device=local.device)
        send[: hi - lo] = local
        gathered = torch.empty(slot * world, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(gathered, send, group=self.group)
        # drop the per-rank padding: rank r contributed shard_bounds(n, world, r) rows
        parts = []
        for r in range(world):
            a, b = shard_bounds(n, world, r)
            parts.append(gathered[r * slot : r * slot + (b - a)])
        return torch.cat(parts)


def argmax_pairs(values_local: torch.Tensor, offset: int, group: Optional[dist.ProcessGroup] = None) -> Tuple[float, int]:
    """Cheaper exchange for argmax-only consumers: all-gather of one (value, global index) pair per rank."""
    if values_local.numel():
        i = int(torch.argmax(values_local))
        pair = torch.tensor([float(values_local[i]), float(offset + i)], dtype=torch.float64, device=values_local.device)
    else:
        pair = torch.tensor([float("-inf"), -1.0], dtype=torch.float64, device=values_local.device)
    if not (dist.is_available() and dist.is_initialized()):
        return float(pair[0]), int(pair[1])
    world = dist.get_world_size(group)
    out = torch.empty(2 * world, dtype=torch.float64, device=pair.device)
    dist.all_gather_into_tensor(out, pair, group=group)
    out = out.reshape(world, 2)
    r = int(torch.argmax(out[:, 0]))
    return float(out[r, 0]), int(out[r, 1])
