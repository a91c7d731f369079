"""Hellinger kernel-kernel: drop-in for bosot/kernels/kernel_kernel_hellinger.py.

Same class name, constructor keywords, attributes (``lengthscale``, ``variance``, ``num_param_samples``,
``num_virtual_points``, ``virtual_X``, ``jitter``) and methods (``transform_X``, ``K``, ``K_diag``,
``get_hellinger_distance_matrix``, ``set_virtual_x_from_dataset``); the arithmetic runs on the GPU through
``bosot_b200.hellinger``.  Deliberate differences:
  * ``transform_X`` returns one ``HellingerOperand`` (flat trees + their device-resident Gram store, built on
    first use) instead of a list of ``EvaluatedKernelWrapper``; ``K`` also takes expression lists / ``FlatTrees``;
  * there is no per-identifier cache: re-sampling a kernel costs microseconds here, the reference's caches
    (:60-102) exist because it costs seconds there.  Results do not depend on call history either way;
  * outputs are float64 torch CUDA tensors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .. import hellinger, sot
from ..encoding import Grammar, encode
from .base_object_kernel import BaseObjectKernel
from .kernel_kernel_grammar_tree import FlatTrees, Parameter


class HellingerOperand:
    """Encoded expressions + (lazily) their Gram store."""

    def __init__(self, flat: FlatTrees):
        self.flat = flat
        self._store = None
        self._space = None

    def __len__(self):
        return len(self.flat)

    def store(self, space: hellinger.HellingerSpace) -> hellinger.GramStore:
        if self._store is None or self._space is not space:
            self._store, self._space = hellinger.grams(self.flat.trees, space), space
        return self._store


class KernelKernelHellinger(BaseObjectKernel):
    def __init__(
        self,
        input_dimension: int,
        base_lengthscale: float,
        base_variance: float,
        num_param_samples: int,
        num_virtual_points: int,
        use_sobol_virtual_points: bool,
        use_hyperprior: bool,
        lengthscale_prior_parameters: Tuple[float, float],
        variance_prior_parameters: Tuple[float, float],
        device: Optional[str] = None,
        **kwargs
    ):
        self.input_dimension = int(input_dimension)
        self.lengthscale = Parameter(base_lengthscale, transform="positive")
        self.variance = Parameter(base_variance, transform="positive")
        # LogNormal(mu, sigma) hyper-priors on the two parameters (reference :118-122), used by the ML-II fit
        self.hyperpriors = {}
        if use_hyperprior:
            self.hyperpriors = {"lengthscale": tuple(lengthscale_prior_parameters), "variance": tuple(variance_prior_parameters)}
        self.num_param_samples = int(num_param_samples)
        self.num_virtual_points = int(num_virtual_points)
        self.virtual_X = hellinger.sobol_points(self.input_dimension, self.num_virtual_points) if use_sobol_virtual_points else None
        self.jitter = 1e-6
        self.device = torch.device(device) if device is not None else None
        self._spaces = {}

    def __deepcopy__(self, memo):
        """ObjectGpModel owns a copy of its kernel (gp_model.py:72): parameters are copied, the virtual points and
        nothing device-resident."""
        import copy

        new = copy.copy(self)
        new.lengthscale, new.variance = copy.deepcopy(self.lengthscale, memo), copy.deepcopy(self.variance, memo)
        new.hyperpriors = dict(self.hyperpriors)
        new._spaces = {}
        return new

    # ---- helpers -------------------------------------------------------------------------------
    def _dev(self) -> torch.device:
        if self.device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("bosot_b200: KernelKernelHellinger needs a CUDA device (no CPU fallback)")
            self.device = torch.device("cuda", torch.cuda.current_device())
        return self.device

    def space(self, grammar: Grammar) -> hellinger.HellingerSpace:
        if self.virtual_X is None:
            raise RuntimeError("virtual points not set: call set_virtual_x_from_dataset(x_data) or use_sobol_virtual_points=True")
        key = (grammar.generator_name, grammar.input_dimension, id(self.virtual_X), self.num_param_samples)
        if key not in self._spaces:
            self._spaces = {key: hellinger.HellingerSpace(grammar, self.virtual_X, self.num_param_samples, self._dev())}
        return self._spaces[key]

    def flatten(self, X) -> FlatTrees:
        if isinstance(X, HellingerOperand):
            return X.flat
        if isinstance(X, FlatTrees):
            return X
        if len(X) == 0:
            raise ValueError("empty expression list")
        grammar = Grammar.for_expressions(X)
        return FlatTrees(sot.as_device_trees(encode(X, grammar), self._dev()), grammar)

    def _operand(self, X) -> HellingerOperand:
        return X if isinstance(X, HellingerOperand) else HellingerOperand(self.flatten(X))

    def _pairs(self, X, X2, want):
        a = self._operand(X)
        b = a if X2 is None else self._operand(X2)
        space = self.space(a.flat.grammar)
        return hellinger.pairs(a.store(space), b.store(space), space, float(self.variance), float(self.lengthscale), want=want)

    # ---- reference API ---------------------------------------------------------------------------
    def transform_X(self, X) -> HellingerOperand:
        """Reference :230-247 (Grams and log-dets of every kernel over the hyper-parameter samples)."""
        return self._operand(X)

    def K(self, X, X2=None):
        """variance * exp(-0.5 * d / lengthscale^2) (reference :128-132)."""
        return self._pairs(X, X2, ("K",))["K"]

    def K_diag(self, X):
        return torch.full((len(X),), float(self.variance), dtype=torch.float64, device=self._dev())

    def get_hellinger_distance_matrix(self, X, X2=None):
        """Expected Hellinger distances (reference :138-183)."""
        return self._pairs(X, X2, ("D",))["D"]

    def set_virtual_x_from_dataset(self, x_data):
        """Random subset of the data as virtual points (reference :254-258; same ``np.random`` call)."""
        indexes = np.random.choice(len(x_data), self.num_virtual_points, replace=False)
        self.virtual_X = np.asarray(x_data, dtype=np.float64)[indexes]
        self._spaces = {}

    # ---- what ObjectGpModel needs from a kernel-kernel -----------------------------------------------
    def gp_parameters(self):
        return [("lengthscale", self.lengthscale), ("variance", self.variance)]

    def device_weights(self) -> np.ndarray:
        return np.array([1.0, 0.0, 0.0])

    def fit_planes(self, flat: FlatTrees):
        """K = variance * exp(-(0.5 d) / lengthscale^2): the fit sees one distance plane, 0.5 * d."""
        op = HellingerOperand(flat)
        space = self.space(flat.grammar)
        D = hellinger.pairs(op.store(space), op.store(space), space, 1.0, 1.0, want=("D",))["D"]
        Df = torch.zeros((3,) + tuple(D.shape), dtype=torch.float64, device=D.device)
        Df[0] = 0.5 * D
        return Df, op

    def make_engine(self, model: dict, noise_variance: float, mean_c: float, acq_kind: int, xi: float, ucb_beta: float):
        space = self.space(model["flat"].grammar)
        return HellingerAcquisitionEngine(model["state"].store(space), model["y"], space, float(self.variance), float(self.lengthscale),
                                          noise_variance, mean_c, acq_kind, xi, ucb_beta)


class HellingerAcquisitionEngine:
    """Counterpart of ``sot.AcquisitionEngine`` for the Hellinger kernel-kernel: evaluated-set Gram store, Cholesky
    and current maximum prepared once; ``acquire_device`` = candidate Grams -> K* -> posterior + acquisition."""

    def __init__(self, eval_store, y, space, variance, lengthscale, noise_variance, mean_c, acq_kind, xi, ucb_beta):
        self.store, self.space = eval_store, space
        self.device = eval_store.G.device
        self.variance, self.lengthscale = float(variance), float(lengthscale)
        self.noise_variance, self.mean_c = float(noise_variance), float(mean_c)
        self.acq_kind, self.xi, self.ucb_beta = int(acq_kind), float(xi), float(ucb_beta)
        y = y.to(self.device, torch.float64).reshape(-1)
        self.K_EE = hellinger.pairs(eval_store, eval_store, space, self.variance, self.lengthscale)["K"]
        self.Linv, self.beta = sot.gp_factorize(self.K_EE, self.noise_variance, y - self.mean_c)
        self.mu_data, self.sigma_data, _ = sot.posterior_acq(self.K_EE, self.Linv, self.beta, self.mean_c, self.variance)
        self.current_max = sot.max_f64(self.mu_data)

    def acquire_device(self, trees: torch.Tensor, out=None):
        Kstar = hellinger.cross(trees, self.store, self.space, self.variance, self.lengthscale)["K"]
        mu, sigma, acq = sot.posterior_acq(Kstar, self.Linv, self.beta, self.mean_c, self.variance, self.acq_kind, self.current_max,
                                           self.xi, self.ucb_beta)
        if out is not None:
            for dst, src in zip(out, (mu, sigma, acq)):
                dst.copy_(src)
            return out
        return mu, sigma, acq

    def raise_on_bad_trees(self) -> None:
        """Same hook as ``sot.AcquisitionEngine``: ``hellinger.cross`` checks the per-tree status itself."""
