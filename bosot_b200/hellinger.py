"""Host side of the Hellinger kernel-kernel (C-ABI ``bosot_hellinger_*``).

Reference: bosot/kernels/kernel_kernel_hellinger.py.  What stays on the host is what the reference does
once per kernel *family*, not per kernel: the Sobol points of the hyper-parameter samples and their
images under the prior quantile functions (:196-206).  Coordinate (i, j) of ``tf.math.sobol_sample`` does not
depend on the number of parameters of the tree, so one table ``quantile[prior, sample i, parameter j]``
serves every tree; the kernels pick column j = running parameter number (leaves left to right, GPflow's
``trainable_parameters`` order inside a leaf).  Everything per tree / per pair runs on the device:
``grams`` (Gram matrices + log-determinants per sample) and ``pairs`` (expected distance, K).
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .encoding import Grammar

# elementary kernels by config name (bosot/configs/kernels/*_configs.py); only kernels WITH priors can be
# sampled (the reference calls ``parameter.prior.quantile``, kernel_kernel_hellinger.py:205)
_KIND_BY_NAME = {
    "RBFWithPrior": _lib.HK_RBF,
    "LinearWithPrior": _lib.HK_LINEAR,
    "PeriodicKernelWithPrior": _lib.HK_PERIODIC,
    "RQWithPrior": _lib.HK_RQ,
}
# Gamma(concentration, rate) per parameter slot, slots in trainable_parameters order; defaults of
# configs/kernels/{rbf,linear,periodic,rational_quadratic}_configs.py
DEFAULT_PRIORS = {
    _lib.HK_RBF: ((2.0, 2.0), (2.0, 3.0)),                 # lengthscales, variance
    _lib.HK_LINEAR: ((2.0, 3.0), (2.0, 3.0)),              # offset, variance
    _lib.HK_PERIODIC: ((2.0, 2.0), (2.0, 2.0), (2.0, 3.0)),  # period, lengthscales, variance
    _lib.HK_RQ: ((2.0, 2.0), (2.0, 2.0), (2.0, 3.0)),      # alpha, lengthscales, variance
}


def sobol_points(dim: int, n: int) -> np.ndarray:
    """``tf.math.sobol_sample(dim, n)`` (skip=0): Joe-Kuo direction numbers in Gray-code order without the
    leading all-zero point, i.e. rows 1..n of SciPy's unscrambled generator."""
    from scipy.stats import qmc

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return qmc.Sobol(d=dim, scramble=False).random(n + 1)[1:]


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError("bosot_b200: %s must be a CUDA tensor (the Hellinger kernel-kernel has no CPU path)" % what)


class HellingerSpace:
    """Symbol table + quantile table of one (grammar, virtual points, number of samples) combination."""

    def __init__(self, grammar: Grammar, virtual_X: np.ndarray, num_param_samples: int, device: torch.device,
                 priors: Optional[Dict[int, Sequence[Tuple[float, float]]]] = None):
        from scipy.special import gammaincinv

        X = np.ascontiguousarray(np.asarray(virtual_X, dtype=np.float64))
        if X.ndim != 2 or X.shape[1] != grammar.input_dimension:
            raise ValueError("virtual_X must be [V, input_dimension]")
        if not 1 <= X.shape[0] <= _lib.HELLINGER_MAX_VIRTUAL:
            raise ValueError("1 <= num_virtual_points <= %d" % _lib.HELLINGER_MAX_VIRTUAL)
        if grammar.input_dimension > _lib.HELLINGER_MAX_DIM:
            raise ValueError("input_dimension <= %d" % _lib.HELLINGER_MAX_DIM)
        self.grammar, self.device = grammar, device
        self.n_virtual, self.n_samples = int(X.shape[0]), int(num_param_samples)
        priors = dict(DEFAULT_PRIORS, **(priors or {}))
        sp = _lib.BosotHellingerSpace()
        planes: Dict[Tuple[float, float], int] = {}
        max_params = 0
        for s, name in enumerate(grammar.symbols):
            base, sep, dim = name.partition("_on_")
            if base not in _KIND_BY_NAME:
                raise ValueError("elementary kernel %r has no hyper-prior to sample from" % name)
            kind = _KIND_BY_NAME[base]
            sp.sym_kind[s] = kind
            sp.sym_dim[s] = int(dim) if sep else -1
            nd = 1 if sep else grammar.input_dimension
            max_params = max(max_params, (1 if kind == _lib.HK_RBF else 2) + (0 if kind == _lib.HK_LINEAR else nd))
        for kind, slots in priors.items():
            for k, ab in enumerate(slots):
                ab = (float(ab[0]), float(ab[1]))
                sp.prior_of[kind][k] = planes.setdefault(ab, len(planes))
        sp.n_symbols, sp.input_dimension = grammar.n_symbols, grammar.input_dimension
        sp.n_virtual, sp.n_samples = self.n_virtual, self.n_samples
        sp.n_sobol_dims, sp.n_priors = _lib.MAX_LEAVES * max_params, len(planes)
        self.c_struct = sp
        u = sobol_points(sp.n_sobol_dims, self.n_samples)  # [S, J]
        table = np.stack([gammaincinv(a, u) / b for (a, b), _ in sorted(planes.items(), key=lambda kv: kv[1])])
        self.quantiles = torch.from_numpy(np.ascontiguousarray(table)).to(device)
        self.virtual_X = torch.from_numpy(X).to(device)
        lib = _lib.load()
        self.record_doubles = int(lib.bosot_hellinger_record_doubles(self.n_virtual))
        if self.record_doubles <= 0:
            _lib.check(self.record_doubles, "hellinger_record_doubles")

    def store_bytes(self, n_trees: int) -> int:
        return self.n_samples * n_trees * (self.record_doubles + 1) * 8


class GramStore:
    """Per (sample, tree): Gram of the composed kernel on the virtual points (+ jitter) and its log|det|."""

    def __init__(self, G: torch.Tensor, logdet: torch.Tensor):
        self.G, self.logdet = G, logdet

    def __len__(self):
        return int(self.G.shape[1])


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def grams(trees: torch.Tensor, space: HellingerSpace, check: bool = True) -> GramStore:
    """``sample_over_hps`` for every tree (kernel_kernel_hellinger.py:192-212): one launch."""
    _need_cuda(trees, "trees")
    n, dev = int(trees.shape[0]), trees.device
    G = torch.empty((space.n_samples, n, space.record_doubles), dtype=torch.float64, device=dev)
    logdet = torch.empty((space.n_samples, n), dtype=torch.float64, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().bosot_hellinger_grams(trees.data_ptr(), n, C.byref(space.c_struct), space.virtual_X.data_ptr(),
                                               space.quantiles.data_ptr(), G.data_ptr(), logdet.data_ptr(), status.data_ptr(), _stream(dev))
    _lib.check(rc, "hellinger_grams")
    if check:
        bad = torch.nonzero(status)
        if bad.numel():
            i = int(bad[0])
            raise ValueError("bosot_b200: tree %d cannot be sampled (code %d)" % (i, int(status[i])))
    return GramStore(G, logdet)


def pairs(a: GramStore, b: GramStore, space: HellingerSpace, variance: float, lengthscale: float,
          want: Sequence[str] = ("K",), out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """Expected Hellinger distances ``D`` and / or ``K = variance * exp(-0.5 D / lengthscale^2)`` between two
    Gram stores, rows = ``a``, columns = ``b`` (kernel_kernel_hellinger.py:128-190).  ``info`` = int64[2] on the
    device: dropped (pair, sample) terms, pairs without a valid sample."""
    na, nb, dev = len(a), len(b), a.G.device
    res = dict(out or {})
    for key in want:
        if key not in ("K", "D"):
            raise ValueError(key)
        if key not in res:
            res[key] = torch.empty((na, nb), dtype=torch.float64, device=dev)
    info = torch.empty(2, dtype=torch.int64, device=dev)
    K, D = res.get("K"), res.get("D")
    ld = (K if K is not None else D).stride(0)
    with torch.cuda.device(dev):
        rc = _lib.load().bosot_hellinger_pairs(a.G.data_ptr(), a.logdet.data_ptr(), na, b.G.data_ptr(), b.logdet.data_ptr(), nb,
                                               space.n_virtual, space.n_samples, float(variance), float(lengthscale),
                                               K.data_ptr() if K is not None else None, D.data_ptr() if D is not None else None, ld,
                                               info.data_ptr(), _stream(dev))
    _lib.check(rc, "hellinger_pairs")
    res["info"] = info
    return res


def cross(trees: torch.Tensor, b: GramStore, space: HellingerSpace, variance: float, lengthscale: float,
          want: Sequence[str] = ("K",), max_store_bytes: int = 2 << 30) -> Dict[str, torch.Tensor]:
    """Rows = many candidate trees, columns = a resident store: candidates are sampled chunk by chunk so that
    their transient Gram store stays below ``max_store_bytes`` (195 KB per tree at S = 100, V = 20)."""
    n = int(trees.shape[0])
    chunk = max(1, min(n, max_store_bytes // max(1, space.store_bytes(1))))
    res = {key: torch.empty((n, len(b)), dtype=torch.float64, device=trees.device) for key in want}
    info = torch.zeros(2, dtype=torch.int64, device=trees.device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        part = pairs(grams(trees[lo:hi], space), b, space, variance, lengthscale, want, out={k: v[lo:hi] for k, v in res.items()})
        info += part["info"]
    res["info"] = info
    return res
