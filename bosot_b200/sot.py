"""Device-side operators of the SOT hot path: thin Python over the C-ABI.

torch is used for device memory, streams and (host side of) the small E x E Cholesky, which
stays with cuSOLVER/cuBLAS through ``torch.linalg``; every per-candidate / per-pair computation
runs in the hand-written sm_100a kernels of ``bosot_b200/csrc`` reached through ``ctypes``.
There is no CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .encoding import Grammar

_DBL3 = C.c_double * 3


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError("bosot_b200: %s must be a CUDA tensor -- the SOT hot path has no CPU implementation" % what)


def as_device_trees(trees, device) -> torch.Tensor:
    """uint8 [N, 64] flat records on ``device`` (accepts numpy / CPU tensors; copies once)."""
    if isinstance(trees, np.ndarray):
        trees = torch.from_numpy(np.ascontiguousarray(trees))
    if trees.dtype != torch.uint8 or trees.dim() != 2 or trees.shape[1] != _lib.TREE_BYTES:
        raise ValueError("flat trees must be uint8 [N, 64]")
    return trees.to(device, non_blocking=True).contiguous()


# ---------------------------------------------------------------------------- featurisation
def featurize(trees: torch.Tensor, grammar: Grammar) -> Dict[str, torch.Tensor]:
    """Ordered sparse histograms of every tree (C-ABI ``bosot_featurize``).  Restates
    ``internal_transform_X`` (bosot/kernels/kernel_kernel_grammar_tree.py:164-185)."""
    _need_cuda(trees, "trees")
    lib = _lib.load()
    n = trees.shape[0]
    dev = trees.device
    o = dict(
        n_nodes=torch.empty(n, dtype=torch.uint8, device=dev),
        n_leaves=torch.empty(n, dtype=torch.uint8, device=dev),
        sub_n=torch.empty(n, dtype=torch.uint8, device=dev),
        sub_id=torch.empty((n, _lib.MAX_NODES), dtype=torch.int64, device=dev),  # bit pattern of uint64 ids
        sub_cnt=torch.empty((n, _lib.MAX_NODES), dtype=torch.uint8, device=dev),
        path_n=torch.empty(n, dtype=torch.uint8, device=dev),
        path_id=torch.empty((n, _lib.MAX_LEAVES), dtype=torch.int16, device=dev),  # bit pattern of uint16 ids
        path_cnt=torch.empty((n, _lib.MAX_LEAVES), dtype=torch.uint8, device=dev),
        sym_cnt=torch.empty((n, grammar.n_symbols), dtype=torch.uint8, device=dev),
        dimvec=torch.empty((n, grammar.n_dim_cols), dtype=torch.float64, device=dev),
        status=torch.empty(n, dtype=torch.int32, device=dev),
    )
    with torch.cuda.device(dev):
        rc = lib.bosot_featurize(
            trees.data_ptr(), n, C.byref(grammar.c_struct),
            o["n_nodes"].data_ptr(), o["n_leaves"].data_ptr(),
            o["sub_n"].data_ptr(), o["sub_id"].data_ptr(), o["sub_cnt"].data_ptr(),
            o["path_n"].data_ptr(), o["path_id"].data_ptr(), o["path_cnt"].data_ptr(),
            o["sym_cnt"].data_ptr(), o["dimvec"].data_ptr(), o["status"].data_ptr(), _stream(dev),
        )
    _lib.check(rc, "featurize")
    return o


def _raise_on_bad_trees(status: torch.Tensor, what: str) -> None:
    bad = torch.nonzero(status)
    if bad.numel():
        i = int(bad[0])
        raise ValueError("bosot_b200: malformed flat tree record in %s at index %d (defect %d)" % (what, i, int(status[i])))


def dense_feature_matrices(fx: Dict[str, torch.Tensor], fx2: Optional[Dict[str, torch.Tensor]] = None):
    """Dense fp64 feature matrices in the reference's column order, per feature type.

    Column index = first-seen order scanning the trees of X then of X2
    (``create_hash_array_mapping``, kernel_kernel_grammar_tree.py:203-211); PATHS / SUBTREES rows
    are ``count / sum(counts)`` and DIM_WISE rows are taken as they are (:198-201, :213-220).
    The device emits each histogram already in dict insertion order; what is left is the
    global first-seen numbering, done here with NumPy on the (small) id streams.
    Returns {feature: (F_X, F_X2 or None)} with feature in {"DIM_WISE", "PATHS", "SUBTREES"}.
    """
    out = {}
    parts = [fx] + ([fx2] if fx2 is not None else [])
    for p in parts:
        _raise_on_bad_trees(p["status"], "feature matrix input")
    dims = [p["dimvec"].cpu().numpy().copy() for p in parts]
    out["DIM_WISE"] = (dims[0], dims[1] if fx2 is not None else None)
    for feat, key_id, key_cnt, key_n in (("PATHS", "path_id", "path_cnt", "path_n"), ("SUBTREES", "sub_id", "sub_cnt", "sub_n")):
        ids = [p[key_id].cpu().numpy() for p in parts]
        cnts = [p[key_cnt].cpu().numpy() for p in parts]
        ns = [p[key_n].cpu().numpy().astype(np.int64) for p in parts]
        width = ids[0].shape[1]
        live = [np.arange(width)[None, :] < n[:, None] for n in ns]
        stream = np.concatenate([i[m] for i, m in zip(ids, live)])  # row-major = tree order, insertion order
        uniq, first = np.unique(stream, return_index=True)
        order = np.argsort(first, kind="stable")
        col_of_sorted = np.empty(len(uniq), dtype=np.int64)
        col_of_sorted[order] = np.arange(len(uniq))
        mats = []
        for i, c, m in zip(ids, cnts, live):
            rows, _ = np.nonzero(m)
            cols = col_of_sorted[np.searchsorted(uniq, i[m])]
            F = np.zeros((i.shape[0], len(uniq)))
            F[rows, cols] = c[m]
            F = F / np.sum(F, axis=1, keepdims=True)
            mats.append(F)
        out[feat] = (mats[0], mats[1] if fx2 is not None else None)
    return out


# ---------------------------------------------------------------------- evaluated-set state
class EvalState:
    """Inverted index over the E evaluated trees (C-ABI ``bosot_eval_build``): replaces the dense
    feature matrix of the training set (kernel_kernel_grammar_tree.py:203-230)."""

    def __init__(self, trees: torch.Tensor, grammar: Grammar):
        _need_cuda(trees, "evaluated trees")
        lib = _lib.load()
        self.grammar = grammar
        self.n_eval = int(trees.shape[0])
        if not (1 <= self.n_eval <= _lib.MAX_EVAL):
            raise ValueError("number of evaluated trees must be in [1, %d]" % _lib.MAX_EVAL)
        self.trees = trees
        nbytes = lib.bosot_eval_state_bytes(self.n_eval)
        self.blob = torch.empty(nbytes, dtype=torch.uint8, device=trees.device)
        status = torch.empty(self.n_eval, dtype=torch.int32, device=trees.device)
        with torch.cuda.device(trees.device):
            rc = lib.bosot_eval_build(trees.data_ptr(), self.n_eval, C.byref(grammar.c_struct), self.blob.data_ptr(), nbytes,
                                      status.data_ptr(), _stream(trees.device))
        _lib.check(rc, "eval_build")
        _raise_on_bad_trees(status, "evaluated set")

    def info(self) -> Dict[str, int]:
        """Diagnostics (C-ABI ``bosot_eval_info``): whether DIM_WISE goes through the block-state tables."""
        h = (C.c_int32 * 4)()
        with torch.cuda.device(self.trees.device):
            rc = _lib.load().bosot_eval_info(self.blob.data_ptr(), self.n_eval, h, _stream(self.trees.device))
        _lib.check(rc, "eval_info")
        return {"dim_wise_states": int(h[0]), "dim_wise_columns": int(h[1]), "max_states_per_block": int(h[2]), "n_eval": int(h[3])}


# ---------------------------------------------------------------------------------- pairs
def sot_pairs(
    trees: torch.Tensor,
    state: EvalState,
    weights: Sequence[float],
    variance: float,
    lengthscale: float,
    want: Sequence[str] = ("K",),
    check: bool = True,
    dtype: torch.dtype = torch.float64,
) -> Dict[str, torch.Tensor]:
    """SOT distances / Gram between candidate trees (rows) and the evaluated set (columns).

    ``weights`` = (w_DIM_WISE, w_PATHS, w_SUBTREES) = alpha_f / sum(alpha)
    (kernel_kernel_grammar_tree.py:159-161).  ``want`` subset of {"K", "D", "Df"}.
    ``dtype=torch.float32``: the reduced-precision variant (C-ABI ``bosot_sot_pairs_f32``)."""
    _need_cuda(trees, "candidate trees")
    if dtype not in (torch.float64, torch.float32):
        raise ValueError("sot_pairs: dtype must be torch.float64 or torch.float32")
    lib = _lib.load()
    n, E, dev = trees.shape[0], state.n_eval, trees.device
    out = {}
    if "K" in want:
        out["K"] = torch.empty((n, E), dtype=dtype, device=dev)
    if "D" in want:
        out["D"] = torch.empty((n, E), dtype=dtype, device=dev)
    if "Df" in want:
        out["Df"] = torch.empty((3, n, E), dtype=dtype, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    scratch = torch.empty(max(16, lib.bosot_sot_pairs_scratch_bytes(n)), dtype=torch.uint8, device=dev)
    entry = lib.bosot_sot_pairs if dtype == torch.float64 else lib.bosot_sot_pairs_f32
    with torch.cuda.device(dev):
        rc = entry(
            trees.data_ptr(), n, C.byref(state.grammar.c_struct), state.blob.data_ptr(), E, _DBL3(*[float(w) for w in weights]),
            float(variance), float(lengthscale), _ptr(out.get("K")), _ptr(out.get("D")), _ptr(out.get("Df")), E,
            status.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream(dev),
        )
    _lib.check(rc, "sot_pairs")
    if check:
        _raise_on_bad_trees(status, "candidates")
    return out


# ------------------------------------------------------------------------- GP factorisation
_INFO = {}


def _pinned_info(dev) -> torch.Tensor:
    """One pinned int32 per device for the status word of bosot_gp_factorize (the kernel stores into it directly)."""
    key = str(dev)
    if key not in _INFO:
        _INFO[key] = torch.zeros(1, dtype=torch.int32).pin_memory()
    return _INFO[key]


def gp_factorize(K_EE: torch.Tensor, noise_variance: float, err: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Small E x E part of the posterior: one fused launch for evaluated sets that fit the shared-memory Cholesky
    (``bosot_gp_factorize``, E <= 160), else left to cuSOLVER/cuBLAS via torch.linalg:
    Lm = chol(K_EE + s^2 I) (GPflow ``base_conditional``), returns
    (LinvF = Lm^-1 packed by ``bosot_pack_linv`` into the zero-padded fragment-major layout of order
    ldl (multiple of 16) that the posterior kernel's fp64 tensor-core tiles read, beta = (K_EE + s^2 I)^-1 err).
    Raises like the reference's TF Cholesky when the matrix is not positive definite."""
    E = K_EE.shape[0]
    lib = _lib.load()
    if E <= lib.bosot_gp_fit_max_eval() and K_EE.stride(1) == 1:
        # one launch (C-ABI bosot_gp_factorize): Cholesky fused with the inversion of the factor in shared memory, LinvF
        # packed and beta written by the same kernel, the status into pinned host memory
        dev = K_EE.device
        ldl = (E + 15) // 16 * 16
        Linv = torch.empty((ldl, ldl), dtype=torch.float64, device=dev)
        beta = torch.empty(E, dtype=torch.float64, device=dev)
        err = err.reshape(E).contiguous()
        info = _pinned_info(dev)
        stream = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.bosot_gp_factorize(K_EE.data_ptr(), E, K_EE.stride(0), float(noise_variance), err.data_ptr(), Linv.data_ptr(), ldl,
                                              beta.data_ptr(), info.data_ptr(), stream.cuda_stream), "gp_factorize")
        stream.synchronize()
        if int(info[0]) != 0:
            raise RuntimeError("Cholesky decomposition was not successful (leading minor %d not positive definite)" % int(info[0]))
        return Linv, beta
    A = K_EE + noise_variance * torch.eye(E, dtype=torch.float64, device=K_EE.device)
    L, info = torch.linalg.cholesky_ex(A)
    if int(info) != 0:
        raise RuntimeError("Cholesky decomposition was not successful (leading minor %d not positive definite)" % int(info))
    eye = torch.eye(E, dtype=torch.float64, device=K_EE.device)
    ldl = (E + 15) // 16 * 16
    Linv_rm = torch.linalg.solve_triangular(L, eye, upper=False).contiguous()
    Linv = torch.empty((ldl, ldl), dtype=torch.float64, device=K_EE.device)
    with torch.cuda.device(K_EE.device):
        _lib.check(_lib.load().bosot_pack_linv(Linv_rm.data_ptr(), E, E, Linv.data_ptr(), ldl, _stream(K_EE.device)), "pack_linv")
    beta = torch.cholesky_solve(err.reshape(E, 1), L).reshape(E).contiguous()
    return Linv, beta


def posterior_acq(
    Kstar: torch.Tensor,
    Linv: torch.Tensor,
    beta: torch.Tensor,
    mean_c: float,
    variance: float,
    acq_kind: int = _lib.ACQ_NONE,
    current_max: Optional[torch.Tensor] = None,
    xi: float = 0.01,
    ucb_beta: float = 3.0,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(mu, sigma, acq) for every row of K* (C-ABI ``bosot_posterior_acq``; a float32 K* takes
    ``bosot_posterior_acq_f32`` and returns float32 tensors)."""
    _need_cuda(Kstar, "K*")
    if Kstar.dtype not in (torch.float64, torch.float32):
        raise ValueError("posterior_acq: K* must be float64 or float32")
    lib = _lib.load()
    n, E = Kstar.shape
    dev = Kstar.device
    mu = torch.empty(n, dtype=Kstar.dtype, device=dev)
    sigma = torch.empty(n, dtype=Kstar.dtype, device=dev)
    acq = torch.empty(n, dtype=Kstar.dtype, device=dev)
    entry = lib.bosot_posterior_acq if Kstar.dtype == torch.float64 else lib.bosot_posterior_acq_f32
    with torch.cuda.device(dev):
        rc = entry(
            Kstar.data_ptr(), n, Kstar.stride(0), E, Linv.data_ptr(), Linv.shape[1], beta.data_ptr(), float(mean_c), float(variance),
            int(acq_kind), _ptr(current_max), float(xi), float(ucb_beta), mu.data_ptr(), sigma.data_ptr(), acq.data_ptr(), _stream(dev),
        )
    _lib.check(rc, "posterior_acq")
    return mu, sigma, acq


def max_f64(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.bosot_max_f64(x.data_ptr(), x.numel(), out.data_ptr(), _stream(x.device)), "max_f64")
    return out


# ------------------------------------------------------------- selection on the device
def topk(values: torch.Tensor, k: int, index_offset: int = 0, want_values: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """The last ``k`` entries of ``np.argsort(values, kind="stable")`` without leaving the device (C-ABI
    ``bosot_topk_f64``): what ``EvolutionaryOptimizerObjects.select`` keeps (evolutionary_optimizer_objects.py:57-64).
    Returns (int64 indices + ``index_offset``, their values), ascending by (value, index); NaN sorts last like NumPy."""
    _need_cuda(values, "values")
    if values.dtype != torch.float64 or values.dim() != 1 or not values.is_contiguous():
        raise ValueError("topk takes a contiguous 1-D float64 tensor")
    lib = _lib.load()
    n, dev = values.numel(), values.device
    k = int(k)
    if not (1 <= k <= n):
        raise ValueError("topk needs 1 <= k <= n")
    idx = torch.empty(k, dtype=torch.int64, device=dev)
    vals = torch.empty(k, dtype=torch.float64, device=dev) if want_values else None
    nbytes = lib.bosot_topk_work_bytes(n, k)
    work = torch.empty(max(16, nbytes), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.bosot_topk_f64(values.data_ptr(), n, k, int(index_offset), idx.data_ptr(), _ptr(vals), work.data_ptr(), work.numel(), _stream(dev))
    _lib.check(rc, "topk")
    return idx, vals


def argmax_first(values: torch.Tensor, index_offset: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``np.argmax`` on the device (C-ABI ``bosot_argmax_f64``): float64 [2] = (max value, its FIRST index + offset).
    ``out`` may be a PINNED HOST tensor: the kernel then stores the pair straight into host memory (unified addressing)
    and the caller only synchronises the stream."""
    _need_cuda(values, "values")
    if values.dtype != torch.float64 or values.dim() != 1 or not values.is_contiguous() or values.numel() < 1:
        raise ValueError("argmax_first takes a non-empty contiguous 1-D float64 tensor")
    lib = _lib.load()
    dev = values.device
    out = out if out is not None else torch.empty(2, dtype=torch.float64, device=dev)
    work = torch.empty(lib.bosot_argmax_work_bytes(), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.bosot_argmax_f64(values.data_ptr(), values.numel(), int(index_offset), out.data_ptr(), work.data_ptr(), work.numel(), _stream(dev))
    _lib.check(rc, "argmax")
    return out


# ------------------------------------------------------------------------ composed engine
class AcquisitionEngine:
    """Everything that depends on the evaluated set and the hyper-parameters, prepared once per
    model state; then ``acquire_*`` turns flat candidate trees into (mu, sigma, acquisition).

    Mirrors what one call of the reference's ``BayesianOptimizerObjects.acquisition_function``
    recomputes from scratch every time (bayesian_optimizer_objects.py:159-168): the E x E Gram,
    its Cholesky, the posterior at the evaluated kernels and ``current_max``.
    """

    def __init__(
        self,
        eval_trees: torch.Tensor,
        y: torch.Tensor,
        grammar: Grammar,
        weights: Sequence[float],
        variance: float,
        lengthscale: float,
        noise_variance: float,
        mean_c: float,
        acq_kind: int = _lib.ACQ_EI,
        xi: float = 0.01,
        ucb_beta: float = 3.0,
        state: Optional[EvalState] = None,
    ):
        _need_cuda(eval_trees, "evaluated trees")
        self.device = eval_trees.device
        self.grammar = grammar
        self.weights = tuple(float(w) for w in weights)
        self.variance, self.lengthscale = float(variance), float(lengthscale)
        self.noise_variance, self.mean_c = float(noise_variance), float(mean_c)
        self.acq_kind, self.xi, self.ucb_beta = int(acq_kind), float(xi), float(ucb_beta)
        self.state = state if state is not None else EvalState(eval_trees, grammar)
        self.n_eval = self.state.n_eval
        y = y.to(self.device, torch.float64).reshape(-1)
        # (the evaluated trees were validated when the state was built: no second status read-back here)
        self.K_EE = sot_pairs(eval_trees, self.state, self.weights, self.variance, self.lengthscale, check=False)["K"]
        self.Linv, self.beta = gp_factorize(self.K_EE, self.noise_variance, y - self.mean_c)
        # posterior at the evaluated kernels -> current_max (bayesian_optimizer_objects.py:164-165)
        self.mu_data, self.sigma_data, _ = posterior_acq(self.K_EE, self.Linv, self.beta, self.mean_c, self.variance)
        self.current_max = max_f64(self.mu_data)
        self._work = None
        self._wdbl = _DBL3(*self.weights)
        # malformed candidate records of the last acquire_* call (their rows are NaN): device count + pinned host copy
        self._bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._h_bad = torch.zeros(1, dtype=torch.int32).pin_memory()

    def _workspace(self, n: int) -> torch.Tensor:
        need = _lib.load().bosot_acquire_work_bytes(n, self.n_eval)
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._work

    def acquire_device(self, trees: torch.Tensor, out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
                       scatter: Optional[Tuple[int, int, int]] = None, dtype: torch.dtype = torch.float64):
        """Candidates already resident in HBM -> (mu, sigma, acq) device tensors.

        ``dtype=torch.float32`` selects the reduced-precision variant (C-ABI ``bosot_acquire_device_f32``: fp32 Gram
        panel and fp32 outputs, fp64 accumulation; 1e-4 relative against the fp64 path).  Never the default.

        ``scatter = (peer_ptrs_dev, n_peers, row_offset)``: the posterior kernel additionally stores every acquisition
        value into the peer-mapped result vectors of all GPUs of the node at ``row_offset + i`` (C-ABI
        ``bosot_acquire_device_scatter``; see ``bosot_b200.distributed.PeerScatter``).

        The call does not synchronise; malformed records poison their own row with NaN and are counted on the device:
        ``bad_tree_count()`` (synchronises) or ``raise_on_bad_trees()`` after the results have been read."""
        _need_cuda(trees, "candidate trees")
        n = trees.shape[0]
        if dtype not in (torch.float64, torch.float32):
            raise ValueError("acquire_device: dtype must be torch.float64 or torch.float32")
        if out is None:
            out = tuple(torch.empty(n, dtype=dtype, device=self.device) for _ in range(3))
        mu, sigma, acq = out
        if any(t.dtype != dtype for t in out):
            raise ValueError("acquire_device: output tensors must have dtype %s" % dtype)
        if dtype == torch.float32 and scatter is not None:
            raise ValueError("acquire_device: the peer scatter is fp64 only")
        w = self._workspace(n)
        lib = _lib.load()
        head = (trees.data_ptr(), n, C.byref(self.grammar.c_struct), self.state.blob.data_ptr(), self.n_eval, self._wdbl,
                self.variance, self.lengthscale, self.Linv.data_ptr(), self.Linv.shape[1], self.beta.data_ptr(), self.mean_c,
                self.acq_kind, self.current_max.data_ptr(), self.xi, self.ucb_beta, w.data_ptr(), w.numel(),
                mu.data_ptr(), sigma.data_ptr(), acq.data_ptr(), self._bad.data_ptr())
        with torch.cuda.device(self.device):
            if dtype == torch.float32:
                rc = lib.bosot_acquire_device_f32(*head, _stream(self.device))
            elif scatter is None:
                rc = lib.bosot_acquire_device(*head, _stream(self.device))
            else:
                ptrs, n_peers, offset = scatter
                rc = lib.bosot_acquire_device_scatter(*head, ptrs, int(n_peers), int(offset), _stream(self.device))
        _lib.check(rc, "acquire_device")
        return mu, sigma, acq

    def bad_tree_count(self) -> int:
        """Malformed candidate records met by the last ``acquire_device`` call (synchronises the stream)."""
        return int(self._bad.item())

    def raise_on_bad_trees(self) -> None:
        n = self.bad_tree_count()
        if n:
            raise ValueError("bosot_b200: %d malformed flat tree record(s) among the candidates (their values are NaN)" % n)

    def acquire_host(self, h_trees: torch.Tensor, h_acq: Optional[torch.Tensor], h_mu: Optional[torch.Tensor] = None,
                     h_sigma: Optional[torch.Tensor] = None, sync: bool = True, scatter: Optional[Tuple[int, int, int]] = None):
        """End-to-end call with HOST buffers (pinned for async copies): H2D of the flat trees,
        both kernels, D2H of the acquisition values (and mu / sigma when given).  With ``scatter`` (see
        ``acquire_device``) the rows are one rank's shard of a pool: the values additionally go to every GPU's
        result vector (C-ABI ``bosot_acquire_host_scatter``) and ``h_acq`` may be None.
        ``sync=True`` waits for the results and raises ``ValueError`` if a candidate record was malformed."""
        if h_trees.is_cuda or (h_acq is not None and h_acq.is_cuda):
            raise ValueError("acquire_host takes host tensors")
        if h_acq is None and scatter is None:
            raise ValueError("acquire_host needs h_acq (or a scatter target)")
        n = h_trees.shape[0]
        w = self._workspace(n)
        head = (h_trees.data_ptr(), n, C.byref(self.grammar.c_struct), self.state.blob.data_ptr(), self.n_eval, self._wdbl,
                self.variance, self.lengthscale, self.Linv.data_ptr(), self.Linv.shape[1], self.beta.data_ptr(), self.mean_c,
                self.acq_kind, self.current_max.data_ptr(), self.xi, self.ucb_beta, w.data_ptr(), w.numel(),
                _ptr(h_mu), _ptr(h_sigma), _ptr(h_acq), self._h_bad.data_ptr())
        with torch.cuda.device(self.device):
            if scatter is None:
                rc = _lib.load().bosot_acquire_host(*head, _stream(self.device))
            else:
                ptrs, n_peers, offset = scatter
                rc = _lib.load().bosot_acquire_host_scatter(*head, ptrs, int(n_peers), int(offset), _stream(self.device))
        _lib.check(rc, "acquire_host")
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
            if int(self._h_bad[0]):
                raise ValueError("bosot_b200: %d malformed flat tree record(s) among the candidates (their values are NaN)"
                                 % int(self._h_bad[0]))
        return h_acq
