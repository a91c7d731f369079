"""Timing of the Hellinger kernel-kernel launches at the C4 shape: 10k candidates x 50 evaluated,
S = 100 hyper-parameter samples, V = 20 virtual points.  Usage: python profiles/tools_hellinger_bench.py [n_cand] [n_eval]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bosot_b200 import hellinger, sot  # noqa: E402
from bosot_b200.encoding import Grammar, random_stress_trees  # noqa: E402

n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_eval = int(sys.argv[2]) if len(sys.argv) > 2 else 50
S, V = 100, 20
dev = torch.device("cuda:0")
gram = Grammar.get("CKSHighDimSearchSpace", 5)
X = np.random.default_rng(0).random((V, 5))
space = hellinger.HellingerSpace(gram, X, S, dev)
ev = sot.as_device_trees(random_stress_trees(n_eval, gram, seed=78), dev)
ca = sot.as_device_trees(random_stress_trees(n_cand, gram, seed=4321), dev)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


st_e = hellinger.grams(ev, space)
st_c = hellinger.grams(ca, space)
ms_g = timed(lambda: hellinger.grams(ca, space, check=False))
out = {"K": torch.empty((n_cand, n_eval), dtype=torch.float64, device=dev)}
ms_p = timed(lambda: hellinger.pairs(st_c, st_e, space, 1.0, 0.5, out=out))
dets = n_cand * n_eval * S
fma = dets * 1330.0
print("grams: %.3f ms (%d trees x %d samples, %.1f M Grams/s)" % (ms_g, n_cand, S, n_cand * S / ms_g / 1e3))
print("pairs: %.3f ms  %.3e determinants/s  useful fp64 %.2f TFLOP/s  (%.3e pairs/s)" % (ms_p, dets / (ms_p * 1e-3), 2 * fma / (ms_p * 1e-3) / 1e12,
                                                                                   n_cand * n_eval / (ms_p * 1e-3)))
t = time.time()
full = hellinger.cross(ca, st_e, space, 1.0, 0.5)
torch.cuda.synchronize()
print("cross (grams + pairs, chunked): %.1f ms; info %s" % ((time.time() - t) * 1e3, full["info"].tolist()))
