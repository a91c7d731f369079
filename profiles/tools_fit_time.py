import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from bosot_b200 import _lib, sot
from bosot_b200.encoding import Grammar, random_stress_trees
lib=_lib.load(); dev=torch.device("cuda:0")
gram=Grammar.get("CKSHighDimSearchSpace",5)
for E in (26, 50, 80, 120, 160):
    ev=sot.as_device_trees(random_stress_trees(E, gram, seed=1), dev)
    st=sot.EvalState(ev, gram)
    Df=sot.sot_pairs(ev, st, np.array([.3,.3,.4]), 1.0, 1.0, want=("Df",))["Df"].contiguous()
    y=torch.randn(E, dtype=torch.float64, device=dev)
    out=torch.empty(9, dtype=torch.float64, device=dev)
    w=(C.c_double*3)(.3,.3,.4); s=torch.cuda.current_stream().cuda_stream
    def k(): lib.bosot_gp_nll_grad(Df.data_ptr(), E, Df.stride(1), y.data_ptr(), w, 1.0, 1.0, 1e-3, 0.0, out.data_ptr(), s)
    for _ in range(5): k()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): k()
    b.record(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(100): k(); out.cpu()
    t1=time.perf_counter()
    print("E=%3d kernel %.1f us   launch+sync+D2H per call %.1f us" % (E, 10*a.elapsed_time(b), 1e4*(t1-t0)))
