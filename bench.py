#!/usr/bin/env python
"""Headline benchmark of the SOT kernel-kernel hot path (BASELINE.json: "SOT OT-pairs/sec and EI
candidates/sec at 1/2/4/8 B200 vs host-CPU ref").

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over the whole candidate pool: flat trees -> in-SM
featurisation -> SOT pair distances against the evaluated set -> Gram -> GP posterior mean/variance
-> Expected Improvement.  When the pool is row-sharded over several GPUs (strong scaling: the pool is
fixed, BASELINE.json config 5) every rank ends with the whole EI vector: by default the posterior kernel
stores its EI values straight into every GPU's vector over NVLink peer memory and the step closes with one
symmetric-memory barrier (--collective fused); --collective nccl uses ONE all-gather after the kernel.

Workload (config.workload): "c5_stress_1M_x_200" = 1,000,000 random grammar trees (depth <= 5,
CKSHighDimSearchSpace d=5) x 200 evaluated trees, hyper-parameters H1 of SURVEY.md 8d.  The C4
headline shape (10,000 candidates x 50 evaluated) is measured in the same run and reported under
"c4_10k_x_50".  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GEN, DIM = "CKSHighDimSearchSpace", 5
POOL, N_EVAL = 1_000_000, 200
C4_POOL, C4_EVAL = 10_000, 50
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback, used only when MEASURED_PEAKS.json is absent
# "trained-like" hyper-parameters H1 of SURVEY.md 8d
H1 = dict(alphas=(0.2, 0.7, 0.4), lengthscale=0.8, variance=1.3, noise_variance=1e-3, mean_c=-0.5)


def algorithmic_bytes_per_pair(n_cand: int, n_eval: int) -> float:
    """one fp64 Gram entry written, each 64-byte flat tree read once."""
    return 8.0 + 64.0 / n_eval + 64.0 / n_cand


class ClockSampler:
    """nvidia-smi SM clocks + throttle reasons sampled DURING the timed regions: one streaming
    ``nvidia-smi -lms 50`` process (a new process per sample would take longer than a whole step)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self._proc, self._t = index, [], None, None

    def _run(self):
        for line in self._proc.stdout:
            cells = [c.strip() for c in line.strip().split(",")]
            if len(cells) >= 2:
                self.rows.append(cells)

    def __enter__(self):
        try:
            self._proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                           "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception:
            self._proc = None
        return self

    def __exit__(self, *a):
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except Exception:
                self._proc.kill()
            self._t.join(timeout=5)

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return dict(json.load(f), source="measured")
    return dict(hbm_gbs=HBM_FALLBACK_GBS, source="fallback")


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, rank: int) -> None:
    if rank != 0:
        return
    from bosot_b200.encoding import Grammar, random_stress_trees
    from oracle import cpu_baseline

    gram = Grammar.get(GEN, DIM)
    ev = random_stress_trees(N_EVAL, gram, seed=77)
    y = np.random.default_rng(4).standard_normal(N_EVAL)
    cores = os.cpu_count() or 1
    sample = 100 * cores  # one reference-shaped call (population 100) per host core per step
    times = []
    for s in range(args.warmup + args.steps):
        ca = random_stress_trees(sample, gram, seed=1234 + s)
        r = cpu_baseline.run(ev, ca, y, H1, GEN, DIM, chunk=100, n_procs=cores)
        if s >= args.warmup:
            times.append(r["seconds"])
    sec = float(np.mean(times))
    value = sample * N_EVAL / sec
    line = dict(
        impl="reference", metric="sot_ot_pairs_per_sec", value=value, unit="pairs/s", n_gpus=args.gpus, steps=args.steps,
        warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
        data="synthetic",
        config=dict(workload="c5_stress_1M_x_200", n_eval=N_EVAL, generator=GEN, input_dimension=DIM, hyper="H1",
                    sample="%d candidates x %d evaluated per step (bounded sample of the 1M pool)" % (sample, N_EVAL)),
        ei_candidates_per_sec=sample / sec,
        cpu_baseline=dict(value=value, unit="pairs/s", cores=cores, kind="port",
                          sample="%d candidates x %d evaluated per step, %d worker processes (1 BLAS thread each) x calls of 100 "
                                 "candidates; K_diag taken as variance" % (sample, N_EVAL, cores),
                          pair_distances_computed_per_call=r["pair_distances_computed_per_call"],
                          pair_distances_credited_per_call=r["pair_distances_credited_per_call"],
                          note="a reference-shaped call recomputes the E x E Gram three times (predict_f at the evaluated "
                               "kernels and at the candidates): only candidates x evaluated pairs are credited"),
        e2e=dict(value=value, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ our arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pool", type=int, default=POOL)
    ap.add_argument("--cpu-sample", type=int, default=0, help="candidates in the cpu_baseline sample (0 = 100 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"],
                    help="N > 1: 'fused' = the posterior kernel stores EI into every GPU's vector over NVLink peer memory + one "
                         "symmetric-memory barrier; 'nccl' = all_gather_into_tensor after the kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from bosot_b200 import _lib, sot
    from bosot_b200.encoding import Grammar, random_stress_trees

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the hot path has no CPU fallback")
    _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.warmup < 3:
        args.warmup = 3  # timing hygiene: at least 3 warm-up steps

    # The clock sampler (one streaming nvidia-smi process) is started BEFORE the inputs are generated: its start-up (NVML
    # initialisation in a second process, ~0.1 s) must be over when the first timed region begins -- started next to it, it
    # landed inside the device-resident timing and cost that figure 4 % (3.83 ms against 3.67 ms for the two launches timed
    # alone, profiles/r3f_bench_n1.json).  From then on it samples every 50 ms across all timed regions.
    clocks = ClockSampler(local_rank)
    clocks.__enter__()

    gram = Grammar.get(GEN, DIM)
    hyp = argparse.Namespace(**H1)
    w = np.asarray(hyp.alphas, dtype=np.float64) / np.sum(np.asarray(hyp.alphas, dtype=np.float64))
    pool = args.pool
    lo, hi = rank * pool // world, (rank + 1) * pool // world
    n_local = hi - lo

    # synthetic inputs: evaluated set replicated, candidate shard per rank
    ev_np = random_stress_trees(N_EVAL, gram, seed=77)
    y = np.random.default_rng(4).standard_normal(N_EVAL)
    ca_np = random_stress_trees(n_local, gram, seed=1234 + rank)
    d_ev = sot.as_device_trees(ev_np, dev)
    d_ca = sot.as_device_trees(ca_np, dev)
    h_ca = torch.from_numpy(ca_np).pin_memory()
    eng = sot.AcquisitionEngine(d_ev, torch.from_numpy(y), gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)

    mu = torch.empty(n_local, dtype=torch.float64, device=dev)
    sigma = torch.empty_like(mu)
    ei_local = torch.empty_like(mu)
    ei_full = torch.empty(pool, dtype=torch.float64, device=dev) if world > 1 else ei_local
    h_ei = torch.empty(pool, dtype=torch.float64).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier(collective=True):
        if world > 1 and collective:
            dist.barrier()
        torch.cuda.synchronize(dev)

    scatter = None
    if world > 1 and args.collective == "fused":
        from bosot_b200.distributed import PeerScatter

        try:
            scatter = PeerScatter(pool, dev)
            assert scatter.bounds == (lo, hi)
            ei_full = scatter.values
        except Exception as exc:  # symmetric memory unavailable on this box: say so and use NCCL
            print("bench: symmetric memory unavailable (%s: %s); falling back to the NCCL all-gather" % (type(exc).__name__, exc),
                  file=sys.stderr)
            scatter = None

    def step_device():
        if scatter is not None:
            eng.acquire_device(d_ca, out=(mu, sigma, ei_local), scatter=scatter.scatter_args())
            scatter.barrier()
            return
        eng.acquire_device(d_ca, out=(mu, sigma, ei_local))
        if world > 1:
            dist.all_gather_into_tensor(ei_full, ei_local)

    def step_e2e():
        # N = 1: host flat trees -> device -> EI -> host through the host-buffer entry point
        eng.acquire_host(h_ca, h_ei, sync=True)

    h_ei_shard = h_ei[lo:hi] if world > 1 else h_ei  # this rank's rows of the result (pinned)
    d_stage = torch.empty_like(d_ca) if (world > 1 and scatter is None) else None

    def step_e2e_multi():
        # N > 1: every rank runs the SAME host-buffer pipeline as N = 1 on its shard (C-ABI bosot_acquire_host_scatter:
        # chunked H2D | kernels | D2H of the shard's EI values into this rank's pinned rows) while the posterior epilogue
        # stores the values into every GPU's vector over NVLink; one barrier closes the step.  Afterwards the host holds
        # the pool's EI values (each rank its rows) and every GPU holds the whole vector.
        if scatter is not None:
            eng.acquire_host(h_ca, h_ei_shard, sync=False, scatter=scatter.scatter_args())
            scatter.barrier()
        else:  # --collective nccl: un-pipelined copy, kernels, all-gather, copy of this rank's rows
            d_stage.copy_(h_ca, non_blocking=True)
            eng.acquire_device(d_stage, out=(mu, sigma, ei_local))
            dist.all_gather_into_tensor(ei_full, ei_local)
            h_ei_shard.copy_(ei_local, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    def timed(fn, steps, warmup, flush_l2=True, collective=True):
        for _ in range(warmup):
            fn()
        barrier(collective)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            if flush_l2:
                flush.zero_()  # untimed: evicts the pool / Gram panel from L2 between timed steps
            a.record()
            fn()
            b.record()
        barrier(collective)
        per_step = [a.elapsed_time(b) for a, b in evs]
        total_ms = sum(per_step)
        timed.last = dict(median_ms=float(np.median(per_step)), min_ms=float(min(per_step)), max_ms=float(max(per_step)))  # this rank's steps
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1 and collective:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / steps

    launches0 = _lib.launch_count()
    ms_step = timed(step_device, args.steps, args.warmup)
    step_stats = dict(timed.last)
    launches = (_lib.launch_count() - launches0) * args.steps // (args.steps + args.warmup)

    # end to end through the host-buffer API (H2D trees, kernels, all-gather, D2H EI inside the timed region)
    ms_e2e = timed(step_e2e_multi if world > 1 else step_e2e, args.steps, args.warmup)

    # per-kernel durations (CUDA events on the launching stream) for the roofline object
    Kbuf = torch.empty((n_local, N_EVAL), dtype=torch.float64, device=dev)
    status = torch.empty(n_local, dtype=torch.int32, device=dev)
    import ctypes as C

    scratch = torch.empty(_lib.load().bosot_sot_pairs_scratch_bytes(n_local), dtype=torch.uint8, device=dev)

    lib = _lib.load()
    wd = (C.c_double * 3)(*[float(x) for x in w])
    stream = torch.cuda.current_stream(dev).cuda_stream

    def k_pairs():
        _lib.check(lib.bosot_sot_pairs(d_ca.data_ptr(), n_local, C.byref(gram.c_struct), eng.state.blob.data_ptr(), N_EVAL, wd,
                                       hyp.variance, hyp.lengthscale, Kbuf.data_ptr(), None, None, N_EVAL, status.data_ptr(),
                                       scratch.data_ptr(), scratch.numel(), stream))

    def k_post():
        _lib.check(lib.bosot_posterior_acq(Kbuf.data_ptr(), n_local, N_EVAL, N_EVAL, eng.Linv.data_ptr(), eng.Linv.shape[1],
                                           eng.beta.data_ptr(), hyp.mean_c, hyp.variance, _lib.ACQ_EI, eng.current_max.data_ptr(),
                                           0.01, 3.0, mu.data_ptr(), sigma.data_ptr(), ei_local.data_ptr(), stream))

    ms_pairs = timed(k_pairs, args.steps, 3, collective=False)
    spread_pairs = dict(timed.last)
    ms_post = timed(k_post, args.steps, 3, collective=False)
    spread_post = dict(timed.last)
    clocks.__exit__()

    # Once per model update (every BO step) the evaluated-set state is rebuilt: inverted index, E x E Gram, Cholesky
    # (cuSOLVER), L^-1 packing, posterior at the evaluated kernels, current_max -- outside `value` and `e2e`, reported here.
    def engine_build_ms(ev_dev, y_np, reps=7):
        y_t = torch.from_numpy(y_np)
        for _ in range(2):
            sot.AcquisitionEngine(ev_dev, y_t, gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
        ev_ms, wall_ms = [], []
        for _ in range(reps):
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            sot.AcquisitionEngine(ev_dev, y_t, gram, w, hyp.variance, hyp.lengthscale, hyp.noise_variance, hyp.mean_c)
            b.record()
            torch.cuda.synchronize(dev)
            wall_ms.append(1e3 * (time.perf_counter() - t0))
            ev_ms.append(a.elapsed_time(b))
        return dict(cuda_event_ms=float(np.median(ev_ms)), wall_ms=float(np.median(wall_ms)))

    build_200 = engine_build_ms(d_ev, y) if rank == 0 else None

    # N > 1: the product path of the scaling run is checked in the run itself.  (1) every rank: the vector the fused
    # scatter left on this GPU == the NCCL all-gather of the per-rank slices, bit for bit; (2) rank 0: 300 random rows
    # of the whole pool against the CPU oracle (the rows' trees are collected from the ranks that own them).
    multi_gpu_parity = None
    if world > 1:
        from bosot_b200.encoding import decode_to_tuple

        eng.acquire_device(d_ca, out=(mu, sigma, ei_local))
        gathered = torch.empty(pool, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, ei_local)
        same = True
        if scatter is not None:
            scatter.values.fill_(float("nan"))
            scatter.barrier()
            step_device()
            torch.cuda.synchronize(dev)
            same = bool(torch.equal(scatter.values, gathered))
            h_ei.fill_(float("nan"))
            step_e2e_multi()
            same = same and bool(torch.equal(scatter.values, gathered)) and bool(torch.equal(h_ei_shard, gathered[lo:hi].cpu()))
        flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        idx = np.sort(np.random.default_rng(99).choice(pool, size=300, replace=False))
        rows = torch.zeros((300, 64), dtype=torch.uint8, device=dev)
        mine = (idx >= lo) & (idx < hi)
        if mine.any():
            rows[torch.from_numpy(np.nonzero(mine)[0]).to(dev)] = d_ca[torch.from_numpy(idx[mine] - lo).to(dev)]
        dist.all_reduce(rows, op=dist.ReduceOp.SUM)
        oracle_ok = True
        if rank == 0:
            from oracle import ref_numpy as rn_p

            ev_t = [decode_to_tuple(r_, gram) for r_ in ev_np]
            ca_t = [decode_to_tuple(r_, gram) for r_ in rows.cpu().numpy()]
            ref = rn_p.acquisition_function(ev_t, y, ca_t, rn_p.Hyper(**H1), GEN, DIM)
            got = gathered[torch.from_numpy(idx).to(dev)].cpu().numpy()
            oracle_ok = bool(np.allclose(got, ref["acq"], rtol=1e-6, atol=1e-12))
        multi_gpu_parity = dict(fused_scatter_equals_all_gather_on_every_rank=bool(int(flag) == 1), oracle_rows_checked=300,
                                oracle_rows_match=oracle_ok, ok=bool(int(flag) == 1) and oracle_ok)

    # fp64 FMA ceiling of this GPU (secondary roofline of the posterior kernel)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    probe_out = torch.empty(sms * 8 * 256, dtype=torch.float64, device=dev)
    iters = 4096

    def k_probe():
        _lib.check(lib.bosot_fp64_fma_probe(probe_out.data_ptr(), sms * 8, 256, iters, stream))

    ms_probe = timed(k_probe, 5, 2, flush_l2=False, collective=False)
    fp64_tflops = 2.0 * sms * 8 * 256 * 8 * iters / (ms_probe * 1e-3) / 1e12

    def k_probe_mma():
        _lib.check(lib.bosot_fp64_mma_probe(probe_out.data_ptr(), sms * 8, 256, iters, stream))

    ms_probe_mma = timed(k_probe_mma, 5, 2, flush_l2=False, collective=False)
    dmma_tflops = 2.0 * sms * 8 * (256 // 32) * 8 * 256 * iters / (ms_probe_mma * 1e-3) / 1e12

    # C4 headline shape: 10k candidates x 50 evaluated (one GPU's worth; measured on every rank, reported from rank 0)
    c4 = {}
    if rank == 0:
        ev4 = random_stress_trees(C4_EVAL, gram, seed=78)
        ca4 = random_stress_trees(C4_POOL, gram, seed=4321)
        y4 = np.random.default_rng(5).standard_normal(C4_EVAL)
        eng4 = sot.AcquisitionEngine(sot.as_device_trees(ev4, dev), torch.from_numpy(y4), gram, w, hyp.variance, hyp.lengthscale,
                                     hyp.noise_variance, hyp.mean_c)
        d4 = sot.as_device_trees(ca4, dev)
        h4 = torch.from_numpy(ca4).pin_memory()
        h4_ei = torch.empty(C4_POOL, dtype=torch.float64).pin_memory()
        o4 = tuple(torch.empty(C4_POOL, dtype=torch.float64, device=dev) for _ in range(3))
        ms4 = timed(lambda: eng4.acquire_device(d4, out=o4), 20, 5, collective=False)
        ms4_e2e = timed(lambda: eng4.acquire_host(h4, h4_ei, sync=True), 20, 5, collective=False)
        alg4 = algorithmic_bytes_per_pair(C4_POOL, C4_EVAL) * C4_POOL * C4_EVAL
        hbm = measured_peaks()["hbm_gbs"]
        build_50 = engine_build_ms(sot.as_device_trees(ev4, dev), y4)
        # the reference's native call size: one EA generation of 100 candidates
        d100 = d4[:100].contiguous()
        o100 = tuple(torch.empty(100, dtype=torch.float64, device=dev) for _ in range(3))
        # device time of one call (L2 flushed between the timed calls like everywhere else: the untimed flush also lets the
        # host enqueue ahead) and the rate of back-to-back calls, which is bound by the host's three launches per call
        ms100 = timed(lambda: eng4.acquire_device(d100, out=o100), 50, 10, collective=False)
        ms100_b2b = timed(lambda: eng4.acquire_device(d100, out=o100), 50, 10, flush_l2=False, collective=False)
        c4 = dict(workload="c4_airfoil_10k_x_50", ms_per_step=ms4, pairs_per_s=C4_POOL * C4_EVAL / (ms4 * 1e-3),
                  ei_candidates_per_s=C4_POOL / (ms4 * 1e-3), e2e_ms=ms4_e2e, e2e_pairs_per_s=C4_POOL * C4_EVAL / (ms4_e2e * 1e-3),
                  e2e_ei_candidates_per_s=C4_POOL / (ms4_e2e * 1e-3),
                  roofline=dict(bound="hbm", achieved=alg4 / (ms4 * 1e-3) / 1e9, peak=hbm, unit="GB/s", frac=alg4 / (ms4 * 1e-3) / 1e9 / hbm,
                                algorithmic_bytes_per_step=alg4, note="whole step (all launches); launch latency, not bandwidth, bounds a pool this small"),
                  engine_build_ms=build_50, e2e_with_update_ms=ms4_e2e + build_50["cuda_event_ms"],
                  call_100_x_50_device_ms=ms100, call_100_x_50_back_to_back_ms=ms100_b2b)

    # Optional fp32 entry points (SURVEY 8b / north_star "1e-4 in fp32") on the headline workload: a separate line with
    # dtype "f32", never the headline.  Gram panel and outputs in fp32, fp64 accumulation in the posterior.
    f32 = {}
    if rank == 0 and world == 1:
        K32 = torch.empty((n_local, N_EVAL), dtype=torch.float32, device=dev)
        o32 = tuple(torch.empty(n_local, dtype=torch.float32, device=dev) for _ in range(3))

        def k_pairs32():
            _lib.check(lib.bosot_sot_pairs_f32(d_ca.data_ptr(), n_local, C.byref(gram.c_struct), eng.state.blob.data_ptr(), N_EVAL, wd,
                                               hyp.variance, hyp.lengthscale, K32.data_ptr(), None, None, N_EVAL, status.data_ptr(),
                                               scratch.data_ptr(), scratch.numel(), stream))

        ms_p32 = timed(k_pairs32, args.steps, 3, collective=False)
        ms_d32 = timed(lambda: eng.acquire_device(d_ca, out=o32, dtype=torch.float32), args.steps, 3, collective=False)
        ref64 = eng.acquire_device(d_ca)
        dev32 = [float(((a.double() - b).abs() / (1e-4 * b.abs() + 1e-4 * hyp.variance ** 0.5)).max()) for a, b in zip(o32, ref64)]
        alg32 = (4.0 + 64.0 / N_EVAL + 64.0 / n_local) * n_local * N_EVAL
        f32 = dict(dtype="f32", workload="c5_stress_1M_x_200", pair_launch_ms=ms_p32, ms_per_step=ms_d32,
                   pairs_per_s=n_local * N_EVAL / (ms_d32 * 1e-3), ei_candidates_per_s=n_local / (ms_d32 * 1e-3),
                   roofline=dict(bound="hbm", kernel="sot_pair_kernel (fp32 outputs)", achieved=alg32 / (ms_p32 * 1e-3) / 1e9,
                                 peak=measured_peaks()["hbm_gbs"], unit="GB/s", frac=alg32 / (ms_p32 * 1e-3) / 1e9 / measured_peaks()["hbm_gbs"],
                                 algorithmic_bytes_per_launch=alg32),
                   share_of_stated_tolerance_used=dict(mu=dev32[0], sigma=dev32[1], ei=dev32[2],
                                                       scale="max |x32 - x64| / (1e-4 |x64| + 1e-4 sqrt(variance)) over the pool; must stay below 1"),
                   note="the fp32 posterior widens K* to fp64 while staging its tiles (generic kernel, no TMA ring): slower than the fp64 step; "
                        "the pair launch is what fp32 speeds up")
        del K32, o32, ref64

    # One BO loop at the shape of BASELINE.json config 4 (Airfoil-shaped: CKSHighDim d = 5, 10k-candidate EA pool per step)
    # through the reference-facing API: model update (ML-II fit on the device) + EA acquisition optimisation per BO step;
    # a deterministic structural objective stands in for the reference's model-evidence oracle (out of scope).
    bo_loop = {}
    if rank == 0 and world == 1:
        try:
            import time as _time

            from bosot_b200.bayesian_optimization.bayesian_optimizer_objects import BayesianOptimizerObjects
            from bosot_b200.configs.bayesian_optimization.bayesian_optimizer_objects_configs import ObjectBOExpectedImprovementEAConfig
            from bosot_b200.configs.kernels.grammar_tree_kernel_kernel_configs import OTWeightedDimsExtendedGrammarKernelConfig
            from bosot_b200.configs.kernels.kernel_grammar_generators.generator_configs import CKSHighDimGeneratorConfig
            from bosot_b200.configs.models.object_gp_model_config import BasicObjectGPModelConfig
            from bosot_b200.kernels.kernel_grammar.generator_factory import GeneratorFactory
            from bosot_b200.models.gp_model import PredictionQuantity
            from bosot_b200.models.model_factory import ModelFactory
            from bosot_b200.models.object_mean_functions import ObjectConstant
            from bosot_b200.oracles.structural_oracle import StructuralOracle

            state = np.random.get_state()
            np.random.seed(0)
            gen_bo = GeneratorFactory.build(CKSHighDimGeneratorConfig(input_dimension=DIM))
            roots = [str(r) for r in gen_bo.search_space.get_root_expressions()]
            model_bo = ModelFactory.build(BasicObjectGPModelConfig(kernel_config=OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=DIM),
                                                                   prediction_quantity=PredictionQuantity.PREDICT_F,
                                                                   perform_multi_start_optimization=False))
            model_bo.set_mean_function(ObjectConstant())
            bo = BayesianOptimizerObjects(**ObjectBOExpectedImprovementEAConfig(n_steps_evolutionary=3, population_evolutionary=C4_POOL).dict())
            bo.set_model(model_bo)
            bo.set_oracle(StructuralOracle(roots, seed=7, noise=0.01))
            bo.set_candidate_generator(gen_bo)
            bo.sample_train_set(3 * len(roots))
            bo.maximize(2)  # warm-up
            torch.cuda.synchronize(dev)
            t0 = _time.perf_counter()
            n_bo = 10
            bo.maximize(n_bo)
            torch.cuda.synchronize(dev)
            bo_loop = dict(workload="c4_airfoil_shape_bo_loop: CKSHighDim d=5, EI + EA with a 10k-candidate pool x 3 generations per step, "
                                    "E = %d..%d evaluated kernels" % (3 * len(roots) + 2, 3 * len(roots) + 2 + n_bo),
                           ms_per_bo_step=1e3 * (_time.perf_counter() - t0) / n_bo, bo_steps=n_bo,
                           ei_evaluations_per_step=3 * C4_POOL, objective="structural oracle (the reference's evidence oracle is out of scope)")
            np.random.set_state(state)
        except Exception as exc:  # the extras never break the headline line
            bo_loop = dict(error="%s: %s" % (type(exc).__name__, exc))

    # Hellinger kernel-kernel at the same 10k x 50 shape: S = 100 hyper-parameter samples, V = 20 virtual points
    hell = {}
    if rank == 0:
        from bosot_b200 import hellinger

        Xv = np.random.default_rng(6).random((20, DIM))
        hsp = hellinger.HellingerSpace(gram, Xv, 100, dev)
        st_e = hellinger.grams(sot.as_device_trees(ev4, dev), hsp)
        st_c = hellinger.grams(d4, hsp)
        Kh = {"K": torch.empty((C4_POOL, C4_EVAL), dtype=torch.float64, device=dev)}
        ms_hg = timed(lambda: hellinger.grams(d4, hsp, check=False), 5, 3, collective=False)
        ms_hp = timed(lambda: hellinger.pairs(st_c, st_e, hsp, 1.0, 0.5, out=Kh), 5, 3, collective=False)
        dets = C4_POOL * C4_EVAL * 100
        hell = dict(workload="hellinger_10k_x_50_S100_V20", gram_store_ms=ms_hg, pair_ms=ms_hp,
                    pairs_per_s=C4_POOL * C4_EVAL / ((ms_hg + ms_hp) * 1e-3), determinants_per_s=dets / (ms_hp * 1e-3),
                    pair_kernel_useful_fp64_tflops=dets * 2 * 1330.0 / (ms_hp * 1e-3) / 1e12,
                    pair_kernel_frac_of_dfma_peak=dets * 2 * 1330.0 / (ms_hp * 1e-3) / 1e12 / fp64_tflops,
                    note="20 x 20 LDL^T = 1330 FMA per determinant; shared-memory / shuffle wavefronts, not the fp64 pipe, bound the pair kernel (DESIGN.md 3.5)")
        if world == 1 and not args.no_cpu_baseline:
            import time as _time

            from bosot_b200.encoding import decode_to_tuple
            from oracle import ref_hellinger as rh

            n_s = 12
            te = [decode_to_tuple(r, gram) for r in ev4]
            tc = [decode_to_tuple(r, gram) for r in ca4[:n_s]]
            t0 = _time.perf_counter()
            Kcpu = rh.K(tc, te, Xv, 100, DIM, 1.0, 0.5)
            dt = _time.perf_counter() - t0
            hell["cpu_port"] = dict(pairs_per_s=n_s * C4_EVAL / dt, cores=1, sample="first %d candidates x %d evaluated in %.1f s (NumPy slogdet)" % (n_s, C4_EVAL, dt),
                                    gpu_matches_cpu_on_sample=bool(np.allclose(Kh["K"][:n_s].cpu().numpy(), Kcpu, rtol=1e-6)))

    peaks = measured_peaks()
    pairs = pool * N_EVAL
    value = pairs / (ms_step * 1e-3)
    e2e_value = pairs / (ms_e2e * 1e-3)
    b_pair = algorithmic_bytes_per_pair(n_local, N_EVAL)
    dominant = "sot_pair_kernel" if ms_pairs >= ms_post else "posterior_kernel"
    if dominant == "sot_pair_kernel":
        alg_bytes = b_pair * n_local * N_EVAL
        dur = ms_pairs
    else:  # posterior: K* panel read once, mu/sigma/EI written
        alg_bytes = (8.0 * N_EVAL + 24.0) * n_local
        dur = ms_post
    traffic = None  # DRAM bytes from the committed ncu capture, only when this run has the same shape
    for tname in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("n_cand") == n_local and tj.get("n_eval") == N_EVAL and dominant in tj:
                traffic = tj[dominant]["dram_bytes"]
                if dominant == "sot_pair_kernel" and "scan_kernel" in tj:
                    traffic += tj["scan_kernel"]["dram_bytes"]  # kernel_ms spans both launches of bosot_sot_pairs
            break
    achieved = alg_bytes / (dur * 1e-3) / 1e9
    post_flops = n_local * (float(N_EVAL) * (N_EVAL + 1) + 4.0 * N_EVAL)  # triangular product + mean + norm
    roofline = dict(
        bound="hbm", kernel=dominant, achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s", frac=achieved / peaks["hbm_gbs"],
        traffic=traffic, peak_source=peaks["source"],
        algorithmic_bytes_per_launch=alg_bytes, kernel_ms=dur,
        kernels_ms=dict(sot_pair_kernel=ms_pairs, posterior_kernel=ms_post),
        kernels_ms_spread=dict(sot_pair_kernel=spread_pairs, posterior_kernel=spread_post),  # rank 0: median / min / max of the timed launches
        kernel_launches="sot_pair_kernel = scan_kernel + sot_pair_kernel (C-ABI bosot_sot_pairs); posterior_kernel = posterior_ws_kernel",
        note="fp64 issue, not HBM, is the practical ceiling of both kernels (DESIGN.md); secondary ceiling below",
        fp64=dict(measured_dfma_tflops=fp64_tflops, measured_dmma_tflops=dmma_tflops,
                  posterior_tflops=post_flops / (ms_post * 1e-3) / 1e12,
                  posterior_frac_of_dmma_peak=post_flops / (ms_post * 1e-3) / 1e12 / dmma_tflops),
    )

    if rank == 0 and world == 1 and not args.no_cpu_baseline and c4:
        # the same CPU port at the headline shape (north_star: >= 100x the reference's CPU kernel-kernel + EI throughput
        # for a 10k-candidate x 50-evaluated matrix): bounded sample, all host cores, doubles as a parity check
        from oracle import cpu_baseline

        cores4 = os.cpu_count() or 1
        sample4 = min(C4_POOL, 200 * cores4)
        r4 = cpu_baseline.run(ev4, ca4[:sample4], y4, H1, GEN, DIM, chunk=100, n_procs=cores4)
        ok4 = bool(np.allclose(o4[2][:sample4].cpu().numpy(), r4["acq"], rtol=1e-6, atol=1e-12))
        c4["cpu_port"] = dict(pairs_per_s=r4["pairs_per_s"], ei_candidates_per_s=r4["cand_per_s"], cores=r4["cores"],
                              sample="first %d candidates x %d evaluated in %.1f s: %d worker processes, calls of 100 candidates; "
                                     "K_diag taken as variance" % (sample4, C4_EVAL, r4["seconds"], r4["cores"]),
                              gpu_matches_cpu_on_sample=ok4)
        c4["e2e_speedup_vs_cpu_port"] = c4["e2e_ei_candidates_per_s"] / r4["cand_per_s"]
        # the reference AS SHIPPED, at its own size: one acquisition_function call on an EA generation of 100 candidates,
        # one process, K_diag through the full N x N x F tensor (kernel_kernel_grammar_tree.py:115-119; SURVEY.md 8d plan 1)
        from oracle import ref_numpy as rn_

        sym4 = rn_.base_kernel_names(GEN, DIM)
        ev_t = [cpu_baseline._decode(r_, sym4) for r_ in ev4]
        ca_t = [cpu_baseline._decode(r_, sym4) for r_ in ca4[:100]]
        t_lit = []
        for _ in range(3):
            t0_ = time.perf_counter()
            lit = rn_.acquisition_function(ev_t, y4, ca_t, rn_.Hyper(**H1), GEN, DIM, literal_kdiag=True)
            t_lit.append(time.perf_counter() - t0_)
        c4["cpu_port"]["literal_call_100_candidates_ms"] = 1e3 * float(np.median(t_lit))
        c4["cpu_port"]["literal_call_matches_gpu"] = bool(np.allclose(o4[2][:100].cpu().numpy(), lit["acq"], rtol=1e-6, atol=1e-12))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline

        cores = os.cpu_count() or 1
        sample = args.cpu_sample or 100 * cores
        r = cpu_baseline.run(ev_np, ca_np[:sample], y, H1, GEN, DIM, chunk=100, n_procs=cores)
        # the sample doubles as a parity check of this very run
        got = ei_local[:sample].cpu().numpy()
        ok = bool(np.allclose(got, r["acq"], rtol=1e-6, atol=1e-12))
        cpu = dict(value=r["pairs_per_s"], unit="pairs/s", cores=r["cores"], kind="port",
                   sample="first %d candidates of the pool x %d evaluated in %.1f s: %d worker processes, calls of 100 "
                          "candidates; K_diag taken as variance" % (sample, N_EVAL, r["seconds"], r["cores"]),
                   ei_candidates_per_sec=r["cand_per_s"], gpu_matches_cpu_on_sample=ok, blas_threads_per_worker=1,
                   pair_distances_computed_per_call=r["pair_distances_computed_per_call"],
                   pair_distances_credited_per_call=r["pair_distances_credited_per_call"])

    if rank == 0:
        line = dict(
            metric="sot_ot_pairs_per_sec", value=value, unit="pairs/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_step, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(workload="c5_stress_1M_x_200", pool=pool, n_eval=N_EVAL, generator=GEN, input_dimension=DIM, hyper="H1",
                        sharding="candidate row blocks, %d per GPU" % n_local, collective=("none" if world == 1 else
                                    "fused: posterior kernel stores EI into every GPU's vector (NVLink peer memory) + 1 symmetric-memory barrier"
                                    if scatter is not None else "all_gather(EI) x1"),
                        l2="flushed between timed steps (512 MB memset, untimed); working set also exceeds L2"),
            ei_candidates_per_sec=pool / (ms_step * 1e-3),
            clocks=clocks.summary(),
            ms_per_step_spread=step_stats,  # rank 0's timed steps: `ms_per_step` is their MEAN (max over ranks), as the contract asks
            e2e=dict(value=e2e_value, unit="pairs/s", ms_per_step=ms_e2e, ei_candidates_per_sec=pool / (ms_e2e * 1e-3),
                     h2d_bytes_per_step=pool * 64, d2h_bytes_per_step=pool * 8,
                     api="AcquisitionEngine.acquire_host (C-ABI bosot_acquire_host)" if world == 1 else
                         ("AcquisitionEngine.acquire_host(scatter=...) per rank on its shard (C-ABI bosot_acquire_host_scatter: H2D of "
                          "the shard | kernels + NVLink scatter | D2H of the shard's EI rows) + 1 symmetric-memory barrier; whole job: "
                          "the pool in, the pool's EI values out (each rank its rows), the full EI vector on every GPU"
                          if scatter is not None else
                          "pinned H2D of the shard + acquire_device + all_gather + pinned D2H of the shard's rows")),
            engine_build_ms=build_200,
            e2e_with_update_ms=(ms_e2e + build_200["cuda_event_ms"]) if build_200 else None,
            multi_gpu_parity=multi_gpu_parity,
            gpu_launches=launches * world,
            roofline=roofline,
            cpu_baseline=cpu,
            c4_10k_x_50=c4,
            hellinger_10k_x_50=hell,
            f32=f32,
            bo_loop_c4=bo_loop,
        )
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
        if multi_gpu_parity is not None and not multi_gpu_parity["ok"]:
            sys.exit("bench: multi-GPU parity check FAILED: %s" % json.dumps(multi_gpu_parity))


if __name__ == "__main__":
    main()
