"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference featuriser.

The reference (boschresearch/bosot, mounted read-only at /root/reference) is pure
Python on top of TensorFlow 2.8 / GPflow 2.4 / TFP 0.16, none of which is
installable in this image.  Its tree classes, search spaces, candidate generators
and the pure-Python/NumPy half of ``OptimalTransportKernelKernel``
(``internal_transform_X`` / ``create_hash_array_mapping`` / ``feature_matrix``,
bosot/kernels/kernel_kernel_grammar_tree.py:164-230) never touch a TF op, so they
run unmodified once the missing third-party modules are replaced by inert stubs.

This module installs those stubs (``install()``) and is used ONLY by
``oracle/make_golden.py`` (fixture generation, in the build container where
/root/reference exists) and by the ``not gpu`` tests that cross-check the
restatement in ``oracle/ref_numpy.py`` against the live reference.  Nothing in the
product (``bosot_b200/``), ``bench.py`` or the ``-m gpu`` tests imports it, and
/root/reference does not exist on the GPU box.

Recipe follows SURVEY.md Appendix B.
"""
from __future__ import annotations

import copy
import math
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("BOSOT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bosot"))


class _Anything:
    """Permissive object: any attribute / call / item yields another _Anything."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, item):
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    """Module whose unknown attributes resolve to permissive stubs."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        return _Anything


def _mod(name: str) -> _StubModule:
    if name in sys.modules and isinstance(sys.modules[name], _StubModule):
        return sys.modules[name]
    m = _StubModule(name)
    m.__path__ = []  # behave as a package so dotted imports work
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_mod(parent), child, m)
    return m


class _Param:
    """Stand-in for gpflow.Parameter: holds a numpy value, nothing else."""

    def __init__(self, value=None, *a, **k):
        self._value = np.asarray(value, dtype=np.float64) if value is not None else None
        self.prior = None
        self.trainable = k.get("trainable", True)

    def numpy(self):
        return self._value

    def assign(self, v):
        self._value = np.asarray(v, dtype=np.float64)


class _Kernel:
    """Stand-in for gpflow.kernels.Kernel: a name, + and * only."""

    def __init__(self, *a, name=None, **k):
        self.name = name

    def __add__(self, other):
        return _Kernel(name="sum")

    def __mul__(self, other):
        return _Kernel(name="product")


class _ParamKernel(_Kernel):
    """Dummy RBF/Polynomial/Periodic/RationalQuadratic exposing assignable priors."""

    def __init__(self, *a, **k):
        super().__init__(name=k.get("name"))
        for attr in ("lengthscales", "variance", "offset", "period", "alpha"):
            setattr(self, attr, _Param(0.0))
        self.base_kernel = a[0] if a and isinstance(a[0], _Kernel) else None


_INSTALLED = False


def install() -> None:
    """Register the stubs.  Idempotent.  Must run before any ``bosot`` import."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    # 1. pydantic v1 API (reference uses BaseSettings with un-annotated overrides)
    import pydantic.v1 as pv1

    sys.modules["pydantic"] = pv1

    # 2. numpy aliases removed in NumPy >= 1.24 / 2.0
    if not hasattr(np, "float"):
        np.float = float  # type: ignore[attr-defined]
    if not hasattr(np, "infty"):
        np.infty = np.inf  # type: ignore[attr-defined]
    if not hasattr(np, "math"):
        np.math = math  # type: ignore[attr-defined]

    # 3. permissive stub packages
    for name in (
        "gpflow",
        "gpflow.kernels",
        "gpflow.utilities",
        "gpflow.utilities.bijectors",
        "gpflow.utilities.ops",
        "gpflow.utilities.traversal",
        "gpflow.config",
        "gpflow.config.__config__",
        "gpflow.ci_utils",
        "gpflow.models",
        "gpflow.models.gpr",
        "gpflow.mean_functions",
        "gpflow.optimizers",
        "gpflow.base",
        "tensorflow",
        "tensorflow.python",
        "tensorflow.python.ops",
        "tensorflow.python.ops.gen_math_ops",
        "tensorflow_probability",
        "tensorflow_probability.distributions",
        "tensorflow_probability.bijectors",
        "gpflux",
        "matplotlib",
        "matplotlib.pyplot",
    ):
        _mod(name)

    # 4. the few real pieces that are subclassed or whose values are used
    gpflow = sys.modules["gpflow"]
    gk = sys.modules["gpflow.kernels"]
    gk.Kernel = _Kernel
    for k in ("RBF", "Polynomial", "Periodic", "RationalQuadratic", "Linear", "Matern52"):
        setattr(gk, k, type(k, (_ParamKernel,), {}))
    gu = sys.modules["gpflow.utilities"]
    gu.to_default_float = lambda x: np.asarray(x, dtype=np.float64)
    gu.deepcopy = copy.deepcopy
    gu.set_trainable = lambda *a, **k: None
    gu.print_summary = lambda *a, **k: None
    gu.reset_cache_bijectors = lambda *a, **k: None
    sys.modules["gpflow.utilities.bijectors"].positive = lambda *a, **k: None
    gc = sys.modules["gpflow.config"]
    gc.set_default_float = lambda *a, **k: None
    gc.set_default_jitter = lambda *a, **k: None
    gc.default_float = lambda: np.float64
    sys.modules["gpflow.config.__config__"].default_float = lambda: np.float64
    sys.modules["gpflow.ci_utils"].ci_niter = lambda n: n
    gpflow.Parameter = _Param
    sys.modules["gpflow.models.gpr"].GPR_with_posterior = type("GPR_with_posterior", (), {})
    gm = sys.modules["gpflow.mean_functions"]
    gm.MeanFunction = type("MeanFunction", (), {"__init__": lambda self, *a, **k: None})
    gm.Constant = type("Constant", (gm.MeanFunction,), {"__init__": lambda self, *a, **k: setattr(self, "c", 0.0)})
    sys.modules["gpflow.base"].Transform = object

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _INSTALLED = True


def hashseed_is_pinned() -> bool:
    """Reference feature keys are Python ``hash()`` of strings (KG:220,269,447-466):
    only equivalence classes and first-seen order are portable across processes, so
    fixtures never store the raw keys; PYTHONHASHSEED only matters for debugging."""
    return os.environ.get("PYTHONHASHSEED", "") not in ("", "random")
