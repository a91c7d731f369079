"""GP over kernel space: drop-in for bosot/models/object_gp_model.py (+ the parts of
bosot/models/gp_model.py and object_gpr.py it inherits: infer / optimize / pertube_parameters /
predictive_dist ...), with GPflow replaced by:

  * the CUDA hot path (``bosot_b200.sot``) for everything per candidate / per pair, and
  * a small fp64 torch routine on the same device for the E x E marginal likelihood of the
    hyper-parameter fit, driven by SciPy L-BFGS-B exactly like
    ``gpflow.optimizers.Scipy`` drives TF (bosot/models/gp_model.py:162-193).

Parameters live in GPflow's unconstrained space during the fit: ``alphas`` through a sigmoid,
``lengthscale`` / ``variance`` through softplus (kernel_kernel_grammar_tree.py:103-107), the
likelihood variance through softplus shifted by GPflow's 1e-6 lower bound, the constant mean as is.
The random perturbation (gp_model.py:263-282) walks the trainable variables in GPflow's order
(kernel.alphas, kernel.lengthscale, kernel.variance, likelihood.variance, mean_function.c) and
consumes the global ``np.random`` stream the same way.
"""
from __future__ import annotations

import contextlib

import copy
import logging
from typing import List, Optional, Tuple

import numpy as np
import torch
from scipy.optimize import minimize
from scipy.stats import norm

from .. import _lib, sot
from ..kernels.base_object_kernel import BaseObjectKernel
from ..kernels.kernel_kernel_grammar_tree import _DEVICE_ORDER, FlatTrees, Parameter
from .gp_model import PredictionQuantity
from .object_mean_functions import MeanFunction, Zero

logger = logging.getLogger(__name__)

LIKELIHOOD_VARIANCE_LOWER_BOUND = 1e-6  # gpflow.likelihoods.Gaussian DEFAULT_VARIANCE_LOWER_BOUND


def _softplus_inv(y: np.ndarray) -> np.ndarray:
    y = np.asarray(y, dtype=np.float64)
    return np.where(y > 30.0, y, np.log(np.expm1(np.minimum(y, 30.0))))


def _softplus(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 30.0, x, np.log1p(np.exp(np.minimum(x, 30.0))))


class ObjectGpModel:
    def __init__(
        self,
        kernel: BaseObjectKernel,
        observation_noise: float,
        optimize_hps: bool,
        train_likelihood_variance: bool,
        pertube_parameters_at_start=False,
        perform_multi_start_optimization=False,
        set_prior_on_observation_noise=False,
        n_starts_for_multistart_opt: int = 5,
        expected_observation_noise: float = 0.1,
        prediction_quantity: PredictionQuantity = PredictionQuantity.PREDICT_Y,
        **kwargs
    ):
        assert isinstance(kernel, BaseObjectKernel)
        self.kernel = copy.deepcopy(kernel)  # the model owns its kernel, like gp_model.py:72
        self._initial_kernel_values = [np.array(p.numpy(), copy=True) for p in self._kernel_parameters()]
        self.observation_noise = observation_noise
        self.optimize_hps = optimize_hps
        self.train_likelihood_variance = train_likelihood_variance
        self.pertube_parameters_at_start = pertube_parameters_at_start
        self.perform_multi_start_optimization = perform_multi_start_optimization
        self.n_starts_for_multistart_opt = n_starts_for_multistart_opt
        self.pertubation_for_multistart_opt = 0.5
        self.set_prior_on_observation_noise = set_prior_on_observation_noise
        self.expected_observation_noise = expected_observation_noise
        self.prediction_quantity = prediction_quantity
        self.use_mean_function = False
        self.mean_function: MeanFunction = Zero()
        self.model = None  # the "built model": evaluated-set state on the device
        self.likelihood_variance = Parameter(np.power(observation_noise, 2.0), trainable=train_likelihood_variance)
        self._engine = None
        self.print_summaries = True
        self.use_fused_fit = True  # bosot_gp_nll_grad; False = torch autograd (the checker / large-E path)

    # ------------------------------------------------------------------ small helpers
    def _kernel_parameters(self) -> List[Parameter]:
        return [p for _, p in self.kernel.gp_parameters()]

    def set_mean_function(self, mean_function: MeanFunction):
        self.use_mean_function = True
        self.mean_function = mean_function

    def set_number_of_restarts(self, n_starts: int):
        self.n_starts_for_multistart_opt = n_starts

    def deactivate_summary_printing(self):
        self.print_summaries = False

    def transform_input(self, x):
        return self.kernel.transform_X(x)

    def reset_model(self):
        """Kernel parameters back to their initial values, built model dropped (gp_model.py:92-98)."""
        if self.model is not None:
            for p, v in zip(self._kernel_parameters(), self._initial_kernel_values):
                p.assign(v)
            self.model = None
            self._engine = None

    # ---------------------------------------------------------------------- build / infer
    def build_model(self, x_data, y_data: np.ndarray):
        """object_gp_model.py:64-76: fixes the data; here that also means the flat encoding of the
        evaluated kernels, their inverted index, and the three E x E per-feature distance matrices
        (constants with respect to every hyper-parameter)."""
        flat = self.kernel.flatten(x_data)
        y = np.asarray(y_data, dtype=np.float64).reshape(-1, 1)
        assert len(flat) == y.shape[0]
        Df, state = self.kernel.fit_planes(flat)
        self.likelihood_variance = Parameter(np.power(self.observation_noise, 2.0), trainable=self.train_likelihood_variance)
        # a mean that depends on the expression (BICMean) is split into its constant c and per-expression offsets: the
        # device path sees targets y - offsets(X) and the constant mean c
        y_eff = y[:, 0] - self._mean_offsets(flat)
        self.model = dict(flat=flat, state=state, Df=Df, y=torch.from_numpy(np.ascontiguousarray(y_eff)).to(flat.trees.device))
        self._engine = None

    def infer(self, x_data, y_data: np.ndarray):
        x_data = self.transform_input(x_data)
        self.build_model(x_data, y_data)
        if self.optimize_hps:
            if self.perform_multi_start_optimization:
                self.multi_start_optimization(self.n_starts_for_multistart_opt, self.pertubation_for_multistart_opt)
            else:
                self.optimize(self.pertube_parameters_at_start)

    # ------------------------------------------------- hyper-parameter fit (ML-II, L-BFGS-B)
    def _trainables(self):
        """(parameter, to_unconstrained, from_unconstrained) in GPflow's variable order."""
        sig_inv = lambda a: np.log(a) - np.log1p(-a)  # noqa: E731
        sig = lambda u: 1.0 / (1.0 + np.exp(-u))  # noqa: E731
        out = []
        for name, p in self.kernel.gp_parameters():
            if p.trainable:
                out.append((p, sig_inv, sig) if name == "alphas" else (p, _softplus_inv, _softplus))
        if self.likelihood_variance.trainable:
            lo = LIKELIHOOD_VARIANCE_LOWER_BOUND
            out.append((self.likelihood_variance, lambda v: _softplus_inv(v - lo), lambda u: _softplus(u) + lo))
        c = getattr(self.mean_function, "c", None)
        if self.use_mean_function and isinstance(c, Parameter) and c.trainable:
            out.append((c, lambda v: np.asarray(v, dtype=np.float64), lambda u: np.asarray(u, dtype=np.float64)))
        return out

    def pertube_parameters(self, factor_bound: float):
        """gp_model.py:263-282."""
        for p, v in zip(self._kernel_parameters(), self._initial_kernel_values):
            p.assign(v)
        if self.train_likelihood_variance:
            self.likelihood_variance.assign(np.power(self.observation_noise, 2.0))
        for p, to_u, from_u in self._trainables():
            u = to_u(p.numpy())
            factor = 1 + np.random.uniform(-1 * factor_bound, factor_bound, size=u.shape)
            if np.isclose(u, 0.0, rtol=1e-07, atol=1e-09).all():
                new_u = (u + np.random.normal(0, 0.05, size=u.shape)) * factor
            else:
                new_u = u * factor
            p.assign(from_u(new_u))

    def _loss_and_grad(self, theta: np.ndarray, layout) -> Tuple[float, np.ndarray]:
        """Negative log marginal likelihood of GPR (+ the exponential prior on the noise when asked) and its
        gradient in unconstrained space.  Fused CUDA kernel (``bosot_gp_nll_grad``: one launch per evaluation,
        analytic gradient) when the evaluated set fits its shared-memory Cholesky, else torch fp64 autograd."""
        Df = self.model["Df"]
        if self.use_fused_fit and Df.is_cuda and Df.shape[1] <= _lib.load().bosot_gp_fit_max_eval():
            return self._loss_and_grad_fused(theta, layout)
        return self._loss_and_grad_autograd(theta, layout)

    def _loss_and_grad_fused(self, theta: np.ndarray, layout) -> Tuple[float, np.ndarray]:
        """One launch of ``bosot_gp_nll_grad`` + the chain rule of the parameter transforms.  Called ~50 times per model
        update, so the host side is plain scalar Python (``math``), not NumPy: the transforms act on at most seven numbers."""
        import ctypes as C
        import math

        k = self.kernel
        th = [float(x) for x in theta]
        vals, pos = {}, 0
        for name, n in layout:
            vals[name] = th[pos : pos + n]
            pos += n
        softplus = lambda x: x if x > 30.0 else math.log1p(math.exp(x))  # noqa: E731
        has_alphas = hasattr(k, "alphas")
        if has_alphas:
            alphas = [1.0 / (1.0 + math.exp(-u)) for u in vals["alphas"]] if "alphas" in vals else [float(a) for a in np.ravel(k.alphas.numpy())]
        ls = softplus(vals["lengthscale"][0]) if "lengthscale" in vals else float(k.lengthscale)
        var = softplus(vals["variance"][0]) if "variance" in vals else float(k.variance)
        lo = LIKELIHOOD_VARIANCE_LOWER_BOUND
        noise = softplus(vals["noise"][0]) + lo if "noise" in vals else float(self.likelihood_variance)
        c = vals["c"][0] if "c" in vals else self.mean_function.value()
        if has_alphas:
            S = math.fsum(alphas)
            order = [_DEVICE_ORDER.index(f) for f in k.feature_type_list]
            w3 = [0.0, 0.0, 0.0]
            for g, a in zip(order, alphas):
                w3[g] += a / S
        else:  # a kernel-kernel with a single distance plane (Hellinger)
            w3 = [float(x) for x in k.device_weights()]
        # The nine results go straight into PINNED HOST memory: under unified addressing the kernel stores to the host
        # pointer itself, so an evaluation is one launch + one stream synchronisation (no device buffer, no copy call).
        if getattr(self, "_fit_out", None) is None:
            self._fit_out = torch.zeros(9, dtype=torch.float64).pin_memory()
            self._fit_out_np = self._fit_out.numpy()
            self._fit_out_ptr = self._fit_out.data_ptr()
        ctx = getattr(self, "_fit_ctx", None)
        if ctx is not None:
            # inside optimize(): the device guard, the stream and the raw pointers were taken once for the whole L-BFGS run
            fn, stream, df_ptr, n_e, ld, y_ptr = ctx
            rc = fn(df_ptr, n_e, ld, y_ptr, (C.c_double * 3)(*w3), ls, var, noise, c, self._fit_out_ptr, stream.cuda_stream)
        else:
            Df, y = self.model["Df"], self.model["y"]
            stream = torch.cuda.current_stream(Df.device)
            with torch.cuda.device(Df.device):
                rc = _lib.load().bosot_gp_nll_grad(Df.data_ptr(), Df.shape[1], Df.stride(1), y.data_ptr(), (C.c_double * 3)(*w3), ls, var, noise, c,
                                                   self._fit_out_ptr, stream.cuda_stream)
        _lib.check(rc, "gp_nll_grad")
        stream.synchronize()
        out = self._fit_out_np.tolist()
        if out[8] != 0.0:
            raise RuntimeError("Cholesky decomposition was not successful (leading minor %d not positive definite)" % int(out[8]))
        loss = out[0]
        d_w3, d_ls, d_var, d_noise, d_c = out[1:4], out[4], out[5], out[6], out[7]
        if self.set_prior_on_observation_noise:  # tfd.Exponential(rate) log-density on the constrained value
            rate = 1.0 / (self.expected_observation_noise * self.expected_observation_noise)
            loss -= math.log(rate) - rate * noise
            d_noise = d_noise + rate
        for pname, (mu, sd) in getattr(k, "hyperpriors", {}).items():  # LogNormal log-density on the constrained value
            x = ls if pname == "lengthscale" else var
            loss -= -math.log(x * sd * math.sqrt(2.0 * math.pi)) - (math.log(x) - mu) ** 2 / (2.0 * sd * sd)
            g = 1.0 / x + (math.log(x) - mu) / (sd * sd * x)
            if pname == "lengthscale":
                d_ls = d_ls + g
            else:
                d_var = d_var + g
        grads = []
        for name, n in layout:
            if name == "alphas":  # w_g = alpha_g / S, alpha = sigmoid(u)
                dot = d_w3[0] * w3[0] + d_w3[1] * w3[1] + d_w3[2] * w3[2]
                grads.extend((d_w3[g] - dot) / S * a * (1.0 - a) for g, a in zip(order, alphas))
            elif name == "lengthscale":
                grads.append(d_ls * (1.0 - math.exp(-ls)))
            elif name == "variance":
                grads.append(d_var * (1.0 - math.exp(-var)))
            elif name == "noise":
                grads.append(d_noise * (1.0 - math.exp(-(noise - lo))))
            else:
                grads.append(d_c)
        return loss, np.array(grads, dtype=np.float64)

    def _loss_and_grad_autograd(self, theta: np.ndarray, layout) -> Tuple[float, np.ndarray]:
        dev = self.model["Df"].device
        t = torch.tensor(theta, dtype=torch.float64, device=dev, requires_grad=True)
        vals = {}
        pos = 0
        for name, n in layout:
            vals[name] = t[pos : pos + n]
            pos += n
        k = self.kernel
        as_t = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)  # noqa: E731
        has_alphas = hasattr(k, "alphas")
        if has_alphas:
            alphas = torch.sigmoid(vals["alphas"]) if "alphas" in vals else as_t(k.alphas.numpy())
        ls = torch.nn.functional.softplus(vals["lengthscale"])[0] if "lengthscale" in vals else as_t(k.lengthscale.numpy())
        var = torch.nn.functional.softplus(vals["variance"])[0] if "variance" in vals else as_t(k.variance.numpy())
        if "noise" in vals:
            noise = torch.nn.functional.softplus(vals["noise"])[0] + LIKELIHOOD_VARIANCE_LOWER_BOUND
        else:
            noise = as_t(self.likelihood_variance.numpy())
        c = vals["c"][0] if "c" in vals else as_t(self.mean_function.value())
        if has_alphas:
            w3 = torch.zeros(3, dtype=torch.float64, device=dev)
            w = alphas / alphas.sum()
            idx = torch.tensor([_DEVICE_ORDER.index(f) for f in k.feature_type_list], device=dev)
            w3 = w3.index_add(0, idx, w)
        else:
            w3 = as_t(k.device_weights())
        D = (w3[:, None, None] * self.model["Df"]).sum(0)
        E = D.shape[0]
        Kmat = var * torch.exp(-D / (ls * ls)) + noise * torch.eye(E, dtype=torch.float64, device=dev)
        L = torch.linalg.cholesky(Kmat)
        err = (self.model["y"] - c).reshape(E, 1)
        a = torch.linalg.solve_triangular(L, err, upper=False)
        lml = -0.5 * (a * a).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * E * np.log(2 * np.pi)
        loss = -lml
        if self.set_prior_on_observation_noise:  # tfd.Exponential(rate) log-density on the constrained value
            rate = 1.0 / np.power(self.expected_observation_noise, 2.0)
            loss = loss - (np.log(rate) - rate * noise)
        for pname, (mu, sd) in getattr(k, "hyperpriors", {}).items():
            x = ls if pname == "lengthscale" else var
            loss = loss - (-torch.log(x * sd * np.sqrt(2.0 * np.pi)) - (torch.log(x) - mu) ** 2 / (2.0 * sd * sd))
        (g,) = torch.autograd.grad(loss, t, allow_unused=True)
        return float(loss.detach()), (g.cpu().numpy() if g is not None else np.zeros_like(theta))

    def training_loss(self) -> float:
        layout, theta, _ = self._pack()
        return self._loss_and_grad(theta, layout)[0]

    def _pack(self):
        names = {id(p): name for name, p in self.kernel.gp_parameters()}
        names[id(self.likelihood_variance)] = "noise"
        layout, chunks, items = [], [], self._trainables()
        for p, to_u, _ in items:
            u = np.atleast_1d(to_u(p.numpy())).astype(np.float64)
            layout.append((names.get(id(p), "c"), len(u)))
            chunks.append(u)
        return layout, (np.concatenate(chunks) if chunks else np.zeros(0)), items

    def _unpack(self, theta, items):
        pos = 0
        for p, _, from_u in items:
            n = int(np.atleast_1d(p.numpy()).size)
            p.assign(from_u(theta[pos : pos + n]))
            pos += n

    def optimize(self, pertube_initial_parameters: bool, pertubation_factor=0.2):
        """gp_model.py:162-193: L-BFGS-B until SciPy reports success; any failure (e.g. Cholesky of a
        non-PD matrix) or non-convergence re-perturbs and retries."""
        if pertube_initial_parameters:
            self.pertube_parameters(pertubation_factor)
        while True:
            success = False
            try:
                layout, theta0, items = self._pack()
                if len(theta0) == 0:
                    return
                with self._fit_context():
                    res = minimize(lambda th: self._loss_and_grad(th, layout), theta0, jac=True, method="L-BFGS-B")
                self._unpack(res.x, items)
                success = bool(res.success)
            except Exception:  # noqa: BLE001 -- mirrors the reference's bare except
                logger.error("Error in optimization - try again")
                self.pertube_parameters(pertubation_factor)
            if success:
                self._engine = None
                return
            logger.warning("Not converged - try again")
            self.pertube_parameters(pertubation_factor)

    @contextlib.contextmanager
    def _fit_context(self):
        """What every loss evaluation of one L-BFGS run needs and none of them changes: the CUDA device guard, the current
        stream, the entry point and the raw pointers of the distance planes and targets (~10 us per evaluation otherwise,
        ~50 evaluations per model update)."""
        Df = None if self.model is None else self.model.get("Df")
        fused = (Df is not None and self.use_fused_fit and Df.is_cuda and Df.shape[1] <= _lib.load().bosot_gp_fit_max_eval())
        if not fused:
            yield
            return
        y = self.model["y"]
        with torch.cuda.device(Df.device):
            self._fit_ctx = (_lib.load().bosot_gp_nll_grad, torch.cuda.current_stream(Df.device), Df.data_ptr(), Df.shape[1], Df.stride(1),
                             y.data_ptr())
            try:
                yield
            finally:
                self._fit_ctx = None

    def multi_start_optimization(self, n_starts: int, pertubation_factor: float):
        """gp_model.py:201-245: best of n_starts perturbed fits by final loss."""
        if self.pertube_parameters_at_start:
            self.pertube_parameters(0.1)
        best = None
        for i in range(n_starts):
            logger.info("Optimization repeat %d/%d", i + 1, n_starts)
            self.optimize(True, pertubation_factor)
            loss = self.training_loss()
            snapshot = [np.array(p.numpy(), copy=True) for p, _, _ in self._trainables()]
            if best is None or loss < best[0]:
                best = (loss, snapshot)
        for (p, _, _), v in zip(self._trainables(), best[1]):
            p.assign(np.maximum(v, 1.000001e-06) if p is self.likelihood_variance else v)
        self._engine = None

    def estimate_model_evidence(self, x_data=None, y_data=None) -> float:
        if self.model is None and x_data is not None and y_data is not None:
            self.infer(x_data, y_data)
        saved, self.set_prior_on_observation_noise = self.set_prior_on_observation_noise, False
        try:
            return -self.training_loss()
        finally:
            self.set_prior_on_observation_noise = saved

    # -------------------------------------------------------------------------- prediction
    def engine(self, acq_kind: int = _lib.ACQ_NONE, xi: float = 0.01, ucb_beta: float = 3.0):
        """Evaluated-set state + Cholesky for the current hyper-parameters (cached until they change)."""
        key = (acq_kind, xi, ucb_beta, tuple(self.kernel.device_weights()), float(self.kernel.variance), float(self.kernel.lengthscale),
               float(self.likelihood_variance), self.mean_function.value())
        if self._engine is None or self._engine[0] != key:
            eng = self.kernel.make_engine(self.model, float(self.likelihood_variance), self.mean_function.value(), acq_kind, xi, ucb_beta)
            self._engine = (key, eng)
        return self._engine[1]

    def _mean_offsets(self, flat) -> np.ndarray:
        off = getattr(self.mean_function, "offsets", None)
        return off(flat) if (self.use_mean_function and off is not None) else np.zeros(len(flat))

    def has_constant_mean(self) -> bool:
        """True when the prior mean is the same for every expression (the fused acquisition path assumes it)."""
        return not (self.use_mean_function and hasattr(self.mean_function, "offsets"))

    def _posterior(self, x_test):
        flat = self.kernel.flatten(self.transform_input(x_test))
        mu, sigma, _ = self.engine().acquire_device(flat.trees)
        if not self.has_constant_mean():
            mu = mu + torch.from_numpy(self._mean_offsets(flat)).to(mu.device)
        return mu, sigma

    def predictive_dist(self, x_test) -> Tuple[np.ndarray, np.ndarray]:
        """(mu[N], sigma[N]) numpy, like gp_model.py:290-306; PREDICT_Y adds the noise variance."""
        mu, sigma = self._posterior(x_test)
        if self.prediction_quantity == PredictionQuantity.PREDICT_Y:
            sigma = torch.sqrt(sigma * sigma + float(self.likelihood_variance))
        return np.squeeze(mu.cpu().numpy()), np.squeeze(sigma.cpu().numpy())

    def predictive_log_likelihood(self, x_test, y_test: np.ndarray) -> np.ndarray:
        mu, sigma = self._posterior(x_test)
        sig_y = torch.sqrt(sigma * sigma + float(self.likelihood_variance)).cpu().numpy()
        return norm.logpdf(np.squeeze(y_test), np.squeeze(mu.cpu().numpy()), np.squeeze(sig_y))

    def entropy_predictive_dist(self, x_test) -> np.ndarray:
        _, sigma = self.predictive_dist(x_test)
        return (0.5 * np.log(2 * np.pi * np.e * np.power(sigma, 2.0))).reshape(-1, 1)
