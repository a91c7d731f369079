"""Row-sharded pool on two GPUs: the fused posterior + peer-memory scatter (C-ABI bosot_acquire_device_scatter,
bosot_b200.distributed.PeerScatter) must give every rank exactly the vector the NCCL all-gather path gives.
Needs two visible GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pool, n_eval, out_dir):
    import torch.distributed as dist

    from bosot_b200 import sot
    from bosot_b200.distributed import PeerScatter, shard_bounds
    from bosot_b200.encoding import Grammar, random_stress_trees

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    gram = Grammar.get("CKSHighDimSearchSpace", 5)
    ev = sot.as_device_trees(random_stress_trees(n_eval, gram, seed=7), dev)
    pool = random_stress_trees(n_pool, gram, seed=8)  # the same pool on every rank
    y = np.random.default_rng(3).standard_normal(n_eval)
    w = np.array([0.2, 0.7, 0.4]) / 1.3
    eng = sot.AcquisitionEngine(ev, torch.from_numpy(y), gram, w, 1.3, 0.8, 1e-3, -0.5)
    lo, hi = shard_bounds(n_pool, world, rank)
    mine = sot.as_device_trees(pool[lo:hi], dev)
    ps = PeerScatter(n_pool, dev)
    for _ in range(3):  # repeated use of the same buffers
        ps.values.fill_(float("nan"))
        ps.barrier()
        mu, sigma, ei = eng.acquire_device(mine, scatter=ps.scatter_args())
        ps.barrier()
        fused = ps.values.clone()
    # reference exchange: padded NCCL all-gather of the local slices
    slot = (n_pool + world - 1) // world
    send = torch.zeros(slot, dtype=torch.float64, device=dev)
    send[: hi - lo] = eng.acquire_device(mine)[2]
    got = torch.empty(slot * world, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(got, send)
    parts = [got[r * slot: r * slot + (shard_bounds(n_pool, world, r)[1] - shard_bounds(n_pool, world, r)[0])] for r in range(world)]
    ref = torch.cat(parts)
    ok = bool(torch.equal(fused, ref)) and bool(torch.equal(fused[lo:hi], ei))
    # single-GPU check of the whole pool on rank 0
    if rank == 0:
        whole = eng.acquire_device(sot.as_device_trees(pool, dev))[2]
        ok = ok and bool(torch.equal(whole, fused))
    with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pool,n_eval", [(20_001, 50), (150_000, 200)])
def test_fused_scatter_equals_all_gather(tmp_path, n_pool, n_eval):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_pool, n_eval, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / ("rank%d.txt" % r)).read() == "ok"
