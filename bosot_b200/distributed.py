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


class PeerScatter:
    """Result vector of a row-sharded pool, replicated on every GPU of the node WITHOUT a collective after the
    kernel: a symmetric-memory allocation (every rank's buffer is peer-mapped on all others over NVLink / NVSwitch)
    that the posterior kernel's epilogue writes directly (C-ABI ``bosot_acquire_device_scatter``), followed by one
    symmetric-memory barrier.  GPU + NCCL process group only; ``ShardedAcquisition`` is the portable path."""

    def __init__(self, n: int, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n = int(n)
        self.values = symm_mem.empty(self.n, dtype=torch.float64, device=device)  # this rank's copy of the full vector
        self.handle = symm_mem.rendezvous(self.values, self.group)
        self.bounds = shard_bounds(self.n, self.world, self.rank)

    def scatter_args(self) -> Tuple[int, int, int]:
        """``scatter=`` argument of ``AcquisitionEngine.acquire_device`` for this rank's shard."""
        return int(self.handle.buffer_ptrs_dev), self.world, self.bounds[0]

    def barrier(self) -> None:
        """All ranks' kernels have stored their shards everywhere once this returns on the current stream."""
        self.handle.barrier()


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
