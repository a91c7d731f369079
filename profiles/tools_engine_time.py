import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from bosot_b200 import _lib, sot
from bosot_b200.encoding import Grammar, random_stress_trees
dev=torch.device("cuda:0"); gram=Grammar.get("CKSHighDimSearchSpace",5)
w=np.array([.2,.7,.4])/1.3
for E in (30, 50, 80):
    ev=sot.as_device_trees(random_stress_trees(E, gram, seed=1), dev); y=torch.randn(E, dtype=torch.float64)
    ca=sot.as_device_trees(random_stress_trees(100, gram, seed=2), dev)
    ca10=sot.as_device_trees(random_stress_trees(10000, gram, seed=3), dev)
    for _ in range(3): eng=sot.AcquisitionEngine(ev, y, gram, w, 1.3, 0.8, 1e-3, -0.5); eng.acquire_device(ca)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(20): eng=sot.AcquisitionEngine(ev, y, gram, w, 1.3, 0.8, 1e-3, -0.5)
    torch.cuda.synchronize(); t1=time.perf_counter()
    for _ in range(20): eng.acquire_device(ca)[2].cpu()
    t2=time.perf_counter()
    for _ in range(20): eng.acquire_device(ca10)[2].cpu()
    t3=time.perf_counter()
    for _ in range(20): st=sot.EvalState(ev, gram)
    torch.cuda.synchronize(); t4=time.perf_counter()
    print("E=%d engine build %.3f ms (EvalState %.3f ms)  acquire 100 + D2H %.3f ms  acquire 10k + D2H %.3f ms" % (E, 50*(t1-t0), 50*(t4-t3), 50*(t2-t1), 50*(t3-t2)))
