"""Host-buffer pipeline tuning: end-to-end time of AcquisitionEngine.acquire_host for per-GPU shard sizes of the scaling
run (1M / 8, / 4, / 2) under different chunking thresholds (C-ABI bosot_host_pipeline_config).
Usage: python profiles/tools_host_pipe.py [n_eval] [reps]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bosot_b200 import _lib, sot  # noqa: E402
from bosot_b200.encoding import Grammar, random_stress_trees  # noqa: E402

n_eval = int(sys.argv[1]) if len(sys.argv) > 1 else 200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 15
dev = torch.device("cuda:0")
lib = _lib.load()
gram = Grammar.get("CKSHighDimSearchSpace", 5)
w = np.array([0.2, 0.7, 0.4]) / 1.3
ev = sot.as_device_trees(random_stress_trees(n_eval, gram, seed=77), dev)
y = np.random.default_rng(4).standard_normal(n_eval)
eng = sot.AcquisitionEngine(ev, torch.from_numpy(y), gram, w, 1.3, 0.8, 1e-3, -0.5)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
full = random_stress_trees(1_000_000, gram, seed=1234)
for pool in (125_000, 250_000, 500_000, 1_000_000):
    h_ca = torch.from_numpy(full[:pool]).pin_memory()
    h_ei = torch.empty(pool, dtype=torch.float64).pin_memory()
    d_ca = h_ca.to(dev)
    out = tuple(torch.empty(pool, dtype=torch.float64, device=dev) for _ in range(3))
    res = []
    for min_cand, first in ((2_000_000, 16384), (100_000, 8192), (100_000, 16384), (100_000, 32768), (100_000, 65536)):
        lib.bosot_host_pipeline_config(min_cand, first)
        ts = []
        for i in range(reps + 3):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.acquire_host(h_ca, h_ei, sync=True)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        res.append("%s/%d: %.4f" % ("off" if min_cand > pool else "on", first, float(np.median(ts))))
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.acquire_device(d_ca, out=out)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    print("pool %7d  device %.4f ms | e2e ms (chunking/first chunk): %s" % (pool, float(np.median(ts)), "  ".join(res)))
