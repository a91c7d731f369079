"""Measures fp64 DFMA / DMMA (mma.sync m8n8k4) throughput as a function of resident warps per SM."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bosot_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0'); sms = torch.cuda.get_device_properties(dev).multi_processor_count
out = torch.empty(sms*32*1024, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
iters = 4096
for fn, name, per_thread in ((lib.bosot_fp64_mma_probe, 'dmma', None), (lib.bosot_fp64_fma_probe, 'dfma', None)):
    for blocks_per_sm, threads in ((1,32),(1,64),(1,128),(1,256),(1,512),(1,1024),(2,1024)):
        nb = sms*blocks_per_sm
        for _ in range(2): fn(out.data_ptr(), nb, threads, iters, st)
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(out.data_ptr(), nb, threads, iters, st); b.record(); torch.cuda.synchronize()
        ms=a.elapsed_time(b)
        if name=='dmma': fl = 2.0*nb*(threads//32)*8*256*iters
        elif name=='dmma2': fl = 2.0*nb*(threads//32)*16*256*iters
        else: fl = 2.0*nb*threads*8*iters
        print(name, 'warps/SM', blocks_per_sm*threads//32, '%.2f TF' % (fl/(ms*1e-3)/1e12))

print("--- register patterns (NA A-fragments x NB B-fragments, which operand is stationary), 8 warps/SM")
pats = {0: (2, 8, "A outer"), 1: (2, 8, "B outer"), 2: (4, 4, "A outer"), 3: (4, 4, "B outer"), 4: (8, 2, "A outer"), 5: (8, 2, "B outer"),
        6: (1, 8, "A outer"), 7: (8, 1, "B outer"), 8: (1, 16, "A outer"), 9: (16, 1, "B outer")}
for pat, (na, nbf, order) in pats.items():
    for threads in (128, 256):
        nb = sms
        arg = (2048 << 4) | pat
        for _ in range(2): lib.bosot_fp64_mma_probe2(out.data_ptr(), nb, threads, arg, st)
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); lib.bosot_fp64_mma_probe2(out.data_ptr(), nb, threads, arg, st); b.record(); torch.cuda.synchronize()
        fl = 2.0*nb*(threads//32)*na*nbf*256*2048
        print("NA=%d NB=%d %s warps/SM %d: %.2f TF" % (na, nbf, order, threads//32, fl/(a.elapsed_time(b)*1e-3)/1e12))

print("--- mixed: even warps DMMA (8 per iteration), odd warps DFMA (64 per iteration per lane)")
for threads in (256, 512, 1024):
    nb = sms
    arg = (2048 << 4) | 15
    for _ in range(2): lib.bosot_fp64_mma_probe2(out.data_ptr(), nb, threads, arg, st)
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record(); lib.bosot_fp64_mma_probe2(out.data_ptr(), nb, threads, arg, st); b.record(); torch.cuda.synchronize()
    w = threads // 32
    fl_t = 2.0*nb*(w//2)*8*256*2048
    fl_f = 2.0*nb*(w//2)*32*64*2048
    ms = a.elapsed_time(b)
    print("warps/SM %d: DMMA part %.2f TF + DFMA part %.2f TF = %.2f TF" % (w, fl_t/(ms*1e-3)/1e12, fl_f/(ms*1e-3)/1e12, (fl_t+fl_f)/(ms*1e-3)/1e12))
