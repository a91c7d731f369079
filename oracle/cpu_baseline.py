"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- times the CPU restatement of the reference's
kernel-kernel + EI path (oracle/ref_numpy.py) on the host cores.

Used by ``bench.py`` for the ``cpu_baseline`` object and for ``--impl reference``.  The reference
itself is pure Python + TF-eager CPU ops and cannot run here (TensorFlow/GPflow are not installable,
and /root/reference does not exist on the GPU box), so the baseline is the oracle *port*: the same
pure-Python featurisation, the same dense [N, F] matrices and [N, M, F] broadcast L1
(bosot/utils/utils.py:168-170), the same posterior and EI.  Two things are kinder to the CPU side
than the literal reference and are stated in the bench output: (1) ``K_diag`` is taken as
``variance`` instead of building the N x N x F tensor (kernel_kernel_grammar_tree.py:115-119), and
(2) the candidate list is cut into calls of ``chunk`` candidates (the reference's EA population
size is 100) that run in parallel worker processes, one per host core -- the reference is single
threaded apart from TF's intra-op pool.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import Dict, List

import numpy as np

from . import ref_numpy as rn

_G: Dict[str, object] = {}


def _decode(record: np.ndarray, symbols: List[str]):
    """Flat pre-order record -> nested tuple (own tiny decoder: the oracle shares no code with the product)."""
    pos = 1

    def rec():
        nonlocal pos
        t = int(record[pos])
        pos += 1
        if t < 0x80:
            return symbols[t]
        left = rec()
        right = rec()
        return ("ADD" if t == 0x80 else "MULTIPLY", left, right)

    return rec()


def _pin_blas_threads(n: int = 1) -> str:
    """One BLAS / OpenMP thread per worker process: with one worker per core, unpinned thread pools oversubscribe the
    host (measured on the GPU box in round 1: 1.31e5 pairs/s unpinned against 2.52e5 with OMP_NUM_THREADS=1)."""
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(n)
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=n)  # pools that were sized when NumPy was imported (before the fork)
        return "threadpoolctl + env"
    except Exception:  # noqa: BLE001
        return "env"


def _init(ev_trees, y, hyper_kw, gen, d):
    _pin_blas_threads(1)
    _G.update(ev=ev_trees, y=y, hyper=rn.Hyper(**hyper_kw), gen=gen, d=d)


def _one_call(cand_trees):
    """One reference-shaped ``acquisition_function`` call (bayesian_optimizer_objects.py:159-168):
    posterior at the evaluated kernels, current_max, posterior at the candidates, EI."""
    r = rn.acquisition_function(_G["ev"], _G["y"], cand_trees, _G["hyper"], _G["gen"], _G["d"], literal_kdiag=False)
    return r["acq"]


def run(eval_records: np.ndarray, cand_records: np.ndarray, y: np.ndarray, hyper_kw: dict, generator_name: str,
        input_dimension: int, chunk: int = 100, n_procs: int = 0) -> dict:
    """Time the port on ``cand_records`` (a bounded sample).  Returns timing + the EI values."""
    symbols = rn.base_kernel_names(generator_name, input_dimension)
    ev = [_decode(r, symbols) for r in eval_records]
    ca = [_decode(r, symbols) for r in cand_records]
    chunks = [ca[i : i + chunk] for i in range(0, len(ca), chunk)]
    n_procs = n_procs or (os.cpu_count() or 1)
    n_procs = max(1, min(n_procs, len(chunks)))
    t0 = time.perf_counter()
    if n_procs == 1:
        _init(ev, y, hyper_kw, generator_name, input_dimension)
        parts = [_one_call(c) for c in chunks]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(n_procs, initializer=_init, initargs=(ev, y, hyper_kw, generator_name, input_dimension)) as pool:
            parts = pool.map(_one_call, chunks, chunksize=1)
    wall = time.perf_counter() - t0
    n, E = len(ca), len(ev)
    # Pair distances one reference-shaped call actually computes (bayesian_optimizer_objects.py:164-166 through
    # GPflow's predict_f): predictive_dist(x_data) = K(E,E) for Kmm + K(E,E) for Kmn (+ K_diag, taken as variance),
    # predictive_dist(x_grid) = K(E,E) for Kmm again + K(E,chunk) -- 3 E^2 + chunk E, of which chunk E are credited.
    per_call = 3 * E * E + min(chunk, n) * E
    return dict(
        pair_distances_computed_per_call=per_call,
        pair_distances_credited_per_call=min(chunk, n) * E,
        blas_threads_per_worker=1 if n_procs > 1 else "unrestricted (single process)",
        seconds=wall,
        n_cand=n,
        n_eval=E,
        cores=n_procs,
        pairs_per_s=n * E / wall,
        cand_per_s=n / wall,
        acq=np.concatenate(parts) if parts else np.zeros(0),
        chunk=chunk,
    )
