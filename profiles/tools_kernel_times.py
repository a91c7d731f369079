"""Quick per-kernel timing at the bench shapes (CUDA events, L2 flushed between launches): scan+pair, posterior.
Usage: python profiles/tools_kernel_times.py [pool] [n_eval] [reps]"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bosot_b200 import _lib, sot  # noqa: E402
from bosot_b200.encoding import Grammar, random_stress_trees  # noqa: E402

pool = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_eval = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
lib = _lib.load()
gram = Grammar.get("CKSHighDimSearchSpace", 5)
alphas = np.array([0.2, 0.7, 0.4])
w = alphas / alphas.sum()
ev = sot.as_device_trees(random_stress_trees(n_eval, gram, seed=77), dev)
import os
_ca_np = random_stress_trees(pool, gram, seed=1234)
if os.environ.get("BOSOT_UNIFORM"):  # experiment: every candidate the same size (no stragglers inside a tile)
    _n = int(os.environ["BOSOT_UNIFORM"])
    _sel = _ca_np[_ca_np[:, 0] == _n]
    _ca_np = _sel[np.arange(pool) % len(_sel)].copy()
ca = sot.as_device_trees(_ca_np, dev)
y = np.random.default_rng(4).standard_normal(n_eval)
eng = sot.AcquisitionEngine(ev, torch.from_numpy(y), gram, w, 1.3, 0.8, 1e-3, -0.5)
K = torch.empty((pool, n_eval), dtype=torch.float64, device=dev)
mu = torch.empty(pool, dtype=torch.float64, device=dev)
sg, ei = torch.empty_like(mu), torch.empty_like(mu)
status = torch.empty(pool, dtype=torch.int32, device=dev)
scratch = torch.empty(lib.bosot_sot_pairs_scratch_bytes(pool), dtype=torch.uint8, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
wd = (C.c_double * 3)(*[float(x) for x in w])
st = torch.cuda.current_stream(dev).cuda_stream


def k_pairs():
    _lib.check(lib.bosot_sot_pairs(ca.data_ptr(), pool, C.byref(gram.c_struct), eng.state.blob.data_ptr(), n_eval, wd, 1.3, 0.8,
                                   K.data_ptr(), None, None, n_eval, status.data_ptr(), scratch.data_ptr(), scratch.numel(), st))


def k_post():
    _lib.check(lib.bosot_posterior_acq(K.data_ptr(), pool, n_eval, n_eval, eng.Linv.data_ptr(), eng.Linv.shape[1], eng.beta.data_ptr(),
                                       -0.5, 1.3, _lib.ACQ_EI, eng.current_max.data_ptr(), 0.01, 3.0, mu.data_ptr(), sg.data_ptr(),
                                       ei.data_ptr(), st))


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


h_ca = ca.cpu().pin_memory()
h_ei = torch.empty(pool, dtype=torch.float64).pin_memory()


def k_e2e():
    eng.acquire_host(h_ca, h_ei, sync=True)


print("info", eng.state.info())
print("pairs  ms median %.4f min %.4f" % timed(k_pairs))
print("post   ms median %.4f min %.4f" % timed(k_post))
print("device ms median %.4f min %.4f" % timed(lambda: eng.acquire_device(ca, out=(mu, sg, ei))))
print("e2e    ms median %.4f min %.4f" % timed(k_e2e))
if os.environ.get("BOSOT_TIME_F32"):  # the optional fp32 entry points (bosot_sot_pairs_f32 / bosot_acquire_device_f32)
    K32 = torch.empty((pool, n_eval), dtype=torch.float32, device=dev)
    o32 = tuple(torch.empty(pool, dtype=torch.float32, device=dev) for _ in range(3))

    def k_pairs32():
        _lib.check(lib.bosot_sot_pairs_f32(ca.data_ptr(), pool, C.byref(gram.c_struct), eng.state.blob.data_ptr(), n_eval, wd, 1.3, 0.8,
                                           K32.data_ptr(), None, None, n_eval, status.data_ptr(), scratch.data_ptr(), scratch.numel(), st))

    print("pairs32  ms median %.4f min %.4f" % timed(k_pairs32))
    print("device32 ms median %.4f min %.4f" % timed(lambda: eng.acquire_device(ca, out=o32, dtype=torch.float32)))
    print("check32  max |ei32 - ei| %.3e" % float((o32[2].double() - ei).abs().max()))
print("check  ei sum %.12e  sigma nan %d" % (float(ei.sum()), int(torch.isnan(sg).sum())))
