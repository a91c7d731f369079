"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy / SciPy fp64) of the reference's alternative
kernel-kernel, the expected Hellinger distance between GP priors.  Not shipped,
never on the product path; only ``tests/`` and ``bench.py``'s CPU-baseline leg import it.

What it restates (reference = boschresearch/bosot, paths relative to /root/reference/bosot):

* ``KernelKernelHellinger.sample_over_hps`` (kernels/kernel_kernel_hellinger.py:192-212): one Sobol
  point per hyper-parameter sample, pushed through the quantile function of every trainable
  parameter's prior, Gram of the composed kernel on the virtual points, jitter, log-determinant;
* ``get_slog_det`` (:214-218), ``hellinger_distance`` (:185-190), ``expected_hellinger_distance``
  (:160-183, including the exclusion of NaN / negative samples), ``get_hellinger_distance_matrix``
  (:138-158) and ``K`` / ``K_diag`` (:128-136);
* the elementary kernels and their Gamma priors (kernels/{rbf,linear,periodic,rational_quadratic}_kernel.py,
  configs/kernels/*_configs.py: lengthscale / period / alpha ~ Gamma(2, 2), variance / offset ~ Gamma(2, 3)),
  the single-dimension slicing of kernels/base_elementary_kernel.py:43-51, and the Sum / Product
  composition of kernels/kernel_grammar/kernel_grammar.py:344-348.

Third-party code the reference calls here and that is ABSENT from this container (environment.yml pins
tensorflow==2.8.0, gpflow==2.4.0, tensorflow-probability==0.16.0); their published behaviour is restated:

* ``tf.math.sobol_sample(dim, n)``: Joe-Kuo direction numbers, Gray-code order, the all-zero point is
  skipped -> rows 1..n of SciPy's unscrambled ``qmc.Sobol``; coordinate j does not depend on ``dim``;
* ``tfd.Gamma(a, b).quantile(u)`` = ``igammainv(a, u) / b`` -> ``scipy.special.gammaincinv``;
* GPflow kernels: ``SquaredExponential`` var*exp(-r2/2), ``RationalQuadratic`` var*(1 + r2/(2 alpha))^-alpha,
  ``Periodic(SquaredExponential)`` var*exp(-sum((sin(pi (x-x')/p)/l)^2)/2), ``Polynomial(degree=1)``
  var*<x,x'> + offset, r2 through ``square_distance`` (the expanded -2xy + x^2 + y^2 form);
  ``Sum`` / ``Product`` flatten nested combinations of their own class and fold left to right;
* ``Module.trainable_parameters`` order (tf.Module._flatten: direct attributes in sorted name order
  first, then sub-modules): leaves left to right; RBF (lengthscales, variance); Linear (offset, variance);
  Periodic (period, lengthscales, variance); RQ (alpha, lengthscales, variance).

Pinning status: PARITY UNPINNED for the sampling half (TF / GPflow / TFP cannot run here and the reference
ships no fixtures).  The pair half (Grams + log-dets -> expected distance -> matrix) is checked against the
live reference methods in ``tests/test_hellinger_oracle.py`` (imported through ``oracle/ref_shim.py``).

Trees are nested tuples like in ``oracle/ref_numpy.py``: leaf = "<ConfigName>[_on_<dim>]", node = (op, l, r).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

JITTER = 1e-6  # kernel_kernel_hellinger.py:126

# (kind, parameter kinds in trainable_parameters order); priors per parameter kind below
_KINDS = {
    "RBFWithPrior": ("rbf", ("lengthscales", "variance")),
    "LinearWithPrior": ("linear", ("offset", "variance")),
    "PeriodicKernelWithPrior": ("periodic", ("period", "lengthscales", "variance")),
    "RQWithPrior": ("rq", ("alpha", "lengthscales", "variance")),
}
# Gamma(concentration, rate): configs/kernels/{rbf,linear,periodic,rational_quadratic}_configs.py
_PRIOR = {"lengthscales": (2.0, 2.0), "variance": (2.0, 3.0), "offset": (2.0, 3.0), "period": (2.0, 2.0), "alpha": (2.0, 2.0)}


def sobol_sample(dim: int, n: int) -> np.ndarray:
    """tf.math.sobol_sample(dim, n) (skip=0) as float64."""
    import warnings

    from scipy.stats import qmc

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # "n should be a power of 2": the reference draws 100 points
        return qmc.Sobol(d=dim, scramble=False).random(n + 1)[1:]


def gamma_quantile(a: float, b: float, u: np.ndarray) -> np.ndarray:
    from scipy.special import gammaincinv

    return gammaincinv(a, u) / b


def leaf_spec(name: str, input_dimension: int):
    """(kind, parameter kinds, active dimension or None, number of active dimensions)."""
    base, sep, dim = name.partition("_on_")
    kind, params = _KINDS[base]
    if sep:
        return kind, params, int(dim), 1
    return kind, params, None, input_dimension


def leaves_of(tree) -> List[str]:
    if isinstance(tree, str):
        return [tree]
    return leaves_of(tree[1]) + leaves_of(tree[2])


def num_parameters(tree, input_dimension: int) -> int:
    total = 0
    for leaf in leaves_of(tree):
        _, params, _, nd = leaf_spec(leaf, input_dimension)
        total += sum(nd if p == "lengthscales" else 1 for p in params)
    return total


def _square_distance(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """gpflow.utilities.ops.square_distance with X2 given (base_elementary_kernel.py:44-45 always passes X2)."""
    return -2.0 * (A @ B.T) + ((A * A).sum(-1)[:, None] + (B * B).sum(-1)[None, :])


def elementary_gram(kind: str, X: np.ndarray, theta: dict) -> np.ndarray:
    if kind == "rbf":
        Xs = X / theta["lengthscales"]
        return theta["variance"] * np.exp(-0.5 * _square_distance(Xs, Xs))
    if kind == "rq":
        Xs = X / theta["lengthscales"]
        return theta["variance"] * (1.0 + 0.5 * _square_distance(Xs, Xs) / theta["alpha"]) ** (-theta["alpha"])
    if kind == "periodic":
        r = np.pi * (X[:, None, :] - X[None, :, :]) / theta["period"]
        return theta["variance"] * np.exp(-0.5 * np.sum(np.square(np.sin(r) / theta["lengthscales"]), -1))
    if kind == "linear":
        return ((X * theta["variance"]) @ X.T + theta["offset"]) ** 1
    raise KeyError(kind)


def gram_of_tree(tree, X: np.ndarray, u_row: np.ndarray, input_dimension: int) -> np.ndarray:
    """Gram of the composed kernel with hyper-parameters = prior quantiles of ``u_row`` (one Sobol point,
    coordinates consumed left to right over the leaves)."""
    cursor = [0]

    def assign(t):  # parameters are assigned in trainable_parameters order: leaves left to right
        if isinstance(t, str):
            kind, params, dim, nd = leaf_spec(t, input_dimension)
            theta = {}
            for p in params:
                k = nd if p == "lengthscales" else 1
                a, b = _PRIOR[p]
                q = gamma_quantile(a, b, u_row[cursor[0] : cursor[0] + k])
                theta[p] = q if p == "lengthscales" else q[0]
                cursor[0] += k
            Xa = X[:, dim : dim + 1] if dim is not None else X
            return ("leaf", elementary_gram(kind, Xa, theta))
        return (t[0], assign(t[1]), assign(t[2]))

    def evaluate(t):
        if t[0] == "leaf":
            return t[1]
        parts = [evaluate(c) for c in _flatten_eval(t, t[0])]
        out = parts[0]
        for p in parts[1:]:
            out = out + p if t[0] == "ADD" else out * p
        return out

    def _flatten_eval(t, op):
        if t[0] != op:
            return [t]
        return _flatten_eval(t[1], op) + _flatten_eval(t[2], op)

    return evaluate(assign(tree))


def slog_det(K: np.ndarray) -> Tuple[np.ndarray, float]:
    Kj = K + np.eye(len(K)) * JITTER
    return Kj, float(np.linalg.slogdet(Kj)[1])


def sample_over_hps(tree, X: np.ndarray, num_param_samples: int, input_dimension: int):
    """-> (grams [S, V, V] with jitter, log-dets [S]) : kernel_kernel_hellinger.py:192-212."""
    u = sobol_sample(num_parameters(tree, input_dimension), num_param_samples)
    grams, dets = [], []
    with np.errstate(all="ignore"):
        for i in range(num_param_samples):
            Kj, ld = slog_det(gram_of_tree(tree, X, u[i], input_dimension))
            grams.append(Kj)
            dets.append(ld)
    return np.stack(grams), np.array(dets)


def hellinger_distance(K1, ld1, K2, ld2) -> float:
    _, ld_avg = slog_det((K1 + K2) / 2.0)
    return 1.0 - math.exp(0.25 * ld1 + 0.25 * ld2 - 0.5 * ld_avg)


def expected_hellinger_distance(g1, d1, g2, d2) -> float:
    total, counter = 0.0, 0
    for i in range(len(d1)):
        with np.errstate(all="ignore"):
            try:
                h = hellinger_distance(g1[i], d1[i], g2[i], d2[i])
            except (OverflowError, ValueError):
                h = float("nan")
        if not math.isnan(h) and h >= 0:
            total += h
            counter += 1
    return total * (1.0 / counter)


def distance_matrix(trees_a: Sequence, trees_b: Optional[Sequence], X: np.ndarray, num_param_samples: int, input_dimension: int) -> np.ndarray:
    ea = [sample_over_hps(t, X, num_param_samples, input_dimension) for t in trees_a]
    if trees_b is None:
        n = len(ea)
        D = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                D[i, j] = D[j, i] = expected_hellinger_distance(*ea[i], *ea[j])
        return D
    eb = [sample_over_hps(t, X, num_param_samples, input_dimension) for t in trees_b]
    return np.array([[expected_hellinger_distance(*a, *b) for b in eb] for a in ea])


def K(trees_a, trees_b, X, num_param_samples, input_dimension, variance: float, lengthscale: float) -> np.ndarray:
    return variance * np.exp(-0.5 * (distance_matrix(trees_a, trees_b, X, num_param_samples, input_dimension) / lengthscale**2))
