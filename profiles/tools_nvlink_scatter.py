"""One process, two GPUs: the posterior kernel on cuda:0 scatters its acquisition values into a vector on cuda:0 AND a
peer-mapped vector on cuda:1 (C-ABI bosot_acquire_device_scatter).  Run plain it checks the bits; run under
  ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:posterior --clock-control none ...
it gives the NVLink traffic of the scatter epilogue (profiles/r2_nvlink_scatter.txt).
Usage: python profiles/tools_nvlink_scatter.py [n_cand] [n_eval]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bosot_b200 import sot  # noqa: E402
from bosot_b200.encoding import Grammar, random_stress_trees  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
n_eval = int(sys.argv[2]) if len(sys.argv) > 2 else 200
assert torch.cuda.device_count() >= 2, "needs two GPUs"
d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
assert torch.cuda.can_device_access_peer(0, 1)
gram = Grammar.get("CKSHighDimSearchSpace", 5)
w = np.array([0.2, 0.7, 0.4]) / 1.3
torch.cuda.set_device(0)
ev = sot.as_device_trees(random_stress_trees(n_eval, gram, seed=77), d0)
ca = sot.as_device_trees(random_stress_trees(n, gram, seed=1234), d0)
y = np.random.default_rng(4).standard_normal(n_eval)
eng = sot.AcquisitionEngine(ev, torch.from_numpy(y), gram, w, 1.3, 0.8, 1e-3, -0.5)
off = 4096
mine = torch.full((off + n,), -1.0, dtype=torch.float64, device=d0)
from bosot_b200 import _lib  # noqa: E402

# The peer vector is a plain cudaMalloc block on cuda:1 (taken from the CUDA runtime directly: torch's caching allocator
# may hand out virtual-memory mappings that cudaDeviceEnablePeerAccess does not cover; the product path uses torch
# symmetric memory, one process per GPU, see bosot_b200/distributed.py).
import ctypes as C  # noqa: E402
import glob  # noqa: E402
import os  # noqa: E402

cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
    glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + \
    glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = C.CDLL(cands[0])
nbytes = 8 * (off + n)
peer_ptr = C.c_void_p()
assert rt.cudaSetDevice(1) == 0
assert rt.cudaMalloc(C.byref(peer_ptr), C.c_size_t(nbytes)) == 0
assert rt.cudaMemset(peer_ptr, 0xFF, C.c_size_t(nbytes)) == 0  # all ones = NaN pattern: "not written"
assert rt.cudaDeviceSynchronize() == 0
assert rt.cudaSetDevice(0) == 0
_lib.check(_lib.load().bosot_enable_peer_access(1), "enable_peer_access")  # cuda:0 maps cuda:1's memory
torch.cuda.synchronize(d0)
table = torch.tensor([mine.data_ptr(), peer_ptr.value], dtype=torch.int64, device=d0)
plain = [t.clone() for t in eng.acquire_device(ca)]
for _ in range(2):
    mu, sg, ei = eng.acquire_device(ca, scatter=(table.data_ptr(), 2, off))
torch.cuda.synchronize(d0)
back = torch.empty(off + n, dtype=torch.float64, device=d0)
assert rt.cudaMemcpy(C.c_void_p(back.data_ptr()), peer_ptr, C.c_size_t(nbytes), 4) == 0  # cudaMemcpyDefault
ok = torch.equal(ei, plain[2]) and torch.equal(mine[off:], ei) and torch.equal(back[off:], ei) and bool(torch.isnan(back[:off]).all())
print("scatter over NVLink: %d candidates x %d evaluated, %d bytes of EI to the peer per launch, bits %s" % (n, n_eval, 8 * n, "OK" if ok else "MISMATCH"))
sys.exit(0 if ok else 1)
