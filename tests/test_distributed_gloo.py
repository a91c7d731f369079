"""CPU, world_size = 2 over gloo: the host-side sharding logic of the N > 1 path (row blocks, padded
all-gather, argmax exchange).  The per-rank compute is a stand-in function of the tree bytes -- the
CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bosot_b200.distributed import ShardedAcquisition, argmax_pairs, padded_shard, shard_bounds
from bosot_b200.encoding import Grammar, random_stress_trees


def test_shard_bounds_partition_every_size():
    for n in (0, 1, 7, 8, 9, 1000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) <= padded_shard(n, world)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _fake_acq(trees: torch.Tensor) -> torch.Tensor:
    w = torch.arange(1, 65, dtype=torch.float64)
    return (trees.to(torch.float64) * w).sum(dim=1) / 1e4


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gram = Grammar.get("CKSHighDimSearchSpace", 5)
        trees = torch.from_numpy(random_stress_trees(n, gram, seed=11))
        full = _fake_acq(trees)
        got = ShardedAcquisition(_fake_acq)(trees)
        assert got.shape == full.shape and torch.equal(got, full)
        lo, hi = shard_bounds(n, world, rank)
        v, i = argmax_pairs(full[lo:hi], lo)
        assert i == int(torch.argmax(full)) and v == float(full.max())
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), got.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 64, 1])
def test_sharded_acquisition_world2(tmp_path, n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b) and a.shape == (n,)


def test_single_process_passthrough():
    t = torch.from_numpy(random_stress_trees(10, Grammar.get("CKSHighDimSearchSpace", 5), seed=1))
    assert torch.equal(ShardedAcquisition(_fake_acq)(t), _fake_acq(t))
    assert argmax_pairs(_fake_acq(t), 5)[1] == 5 + int(torch.argmax(_fake_acq(t)))
