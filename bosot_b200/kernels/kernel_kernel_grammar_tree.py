"""SOT kernel-kernel: drop-in for bosot/kernels/kernel_kernel_grammar_tree.py.

Same class names, constructor keywords, attributes (``alphas``, ``lengthscale``, ``variance``,
``feature_type_list``) and methods; the arithmetic runs on the GPU through ``bosot_b200.sot``.
Differences that are deliberate and documented in DESIGN.md:
  * inputs may be expression objects (anything ``bosot_b200.encoding.encode`` accepts) or an already
    encoded ``FlatTrees``; outputs are float64 torch CUDA tensors where the reference returns tf tensors;
  * ``K_diag`` returns ``variance`` for every input instead of building the full N x N distance
    matrix for its diagonal (reference :115-119) -- the diagonal distance is exactly 0.
"""
from __future__ import annotations

from enum import Enum
from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import sot
from ..encoding import Grammar, encode
from .base_object_kernel import BaseObjectKernel
from .kernel_grammar.kernel_grammar import BaseKernelGrammarExpression, KernelGrammarExpressionTransformer


class FeatureType(Enum):
    SUBTREES = 0
    REDUCED_ELEMENTARY_PATHS = 1
    DIM_WISE_WEIGHTED_ELEMENTARY_COUNT = 2


# order of the weight triple the C-ABI takes
_DEVICE_ORDER = (FeatureType.DIM_WISE_WEIGHTED_ELEMENTARY_COUNT, FeatureType.REDUCED_ELEMENTARY_PATHS, FeatureType.SUBTREES)


class FlatTrees:
    """Encoded expressions resident on the device (+ the grammar they belong to)."""

    def __init__(self, trees: torch.Tensor, grammar: Grammar):
        self.trees, self.grammar = trees, grammar

    def __len__(self):
        return int(self.trees.shape[0])


class Parameter:
    """Minimal stand-in for gpflow.Parameter: a float64 value (numpy), a trainable flag and an
    optional (transform, inverse) pair used by the hyper-parameter fit."""

    def __init__(self, value, trainable=True, transform=None):
        self._value = np.asarray(value, dtype=np.float64)
        self.trainable = bool(trainable)
        self.transform = transform  # "positive" (softplus) | "sigmoid" | None

    def numpy(self):
        return self._value

    def assign(self, value):
        self._value = np.asarray(value, dtype=np.float64).reshape(self._value.shape)

    def __float__(self):
        return float(self._value)

    def __repr__(self):
        return "Parameter(%r, trainable=%r, transform=%r)" % (self._value, self.trainable, self.transform)


class BaseKernelGrammarKernel(BaseObjectKernel):
    def __init__(self, transform_to_normal: bool, **kwargs):
        self.transform_to_normal = transform_to_normal

    def transform_X(self, X):
        if self.transform_to_normal and not isinstance(X, FlatTrees):
            return [KernelGrammarExpressionTransformer.transform_to_normal_form(x) for x in X]
        return X


class OptimalTransportKernelKernel(BaseKernelGrammarKernel):
    def __init__(
        self,
        feature_type_list: List[FeatureType],
        base_variance: float,
        base_lengthscale: float,
        base_alpha: float,
        alpha_trainable: bool,
        parameters_trainable: bool,
        transform_to_normal: bool,
        device: Optional[str] = None,
        **kwargs
    ):
        super().__init__(transform_to_normal)
        self.feature_type_list = list(feature_type_list)
        if not self.feature_type_list or any(f not in _DEVICE_ORDER for f in self.feature_type_list):
            raise ValueError("feature_type_list must hold FeatureType members")
        self.alphas = Parameter(np.repeat(base_alpha, len(self.feature_type_list)), trainable=parameters_trainable and alpha_trainable, transform="sigmoid")
        self.lengthscale = Parameter(base_lengthscale, trainable=parameters_trainable, transform="positive")
        self.variance = Parameter(base_variance, trainable=parameters_trainable, transform="positive")
        self.device = torch.device(device) if device is not None else None

    # ---- helpers -------------------------------------------------------------------------
    def _dev(self) -> torch.device:
        if self.device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("bosot_b200: OptimalTransportKernelKernel needs a CUDA device (no CPU fallback)")
            self.device = torch.device("cuda", torch.cuda.current_device())
        return self.device

    def flatten(self, X) -> FlatTrees:
        """Expression objects -> device-resident flat records (the boundary of the hot path)."""
        if isinstance(X, FlatTrees):
            return X
        if len(X) == 0:
            raise ValueError("empty expression list")
        grammar = Grammar.for_expressions(X)
        return FlatTrees(sot.as_device_trees(encode(X, grammar), self._dev()), grammar)

    def device_weights(self) -> np.ndarray:
        """alpha_f / sum(alpha) (reference :140, :159) laid out as (DIM_WISE, PATHS, SUBTREES); 0 = feature unused."""
        a = self.alphas.numpy()
        w = a / np.sum(a)
        out = np.zeros(3)
        for f, wf in zip(self.feature_type_list, w):
            out[_DEVICE_ORDER.index(f)] += wf
        return out

    # ---- what ObjectGpModel needs from a kernel-kernel ----------------------------------------
    def gp_parameters(self):
        """(name, Parameter) in GPflow's variable order."""
        return [("alphas", self.alphas), ("lengthscale", self.lengthscale), ("variance", self.variance)]

    hyperpriors = {}

    def fit_planes(self, flat: FlatTrees):
        """The three E x E per-feature distance matrices (constants of the hyper-parameter fit) + the evaluated-set state."""
        state = sot.EvalState(flat.trees, flat.grammar)
        return sot.sot_pairs(flat.trees, state, (1.0, 1.0, 1.0), 1.0, 1.0, want=("Df",))["Df"], state

    def make_engine(self, model: dict, noise_variance: float, mean_c: float, acq_kind: int, xi: float, ucb_beta: float):
        return sot.AcquisitionEngine(model["flat"].trees, model["y"], model["flat"].grammar, self.device_weights(), float(self.variance),
                                     float(self.lengthscale), noise_variance, mean_c, acq_kind=acq_kind, xi=xi, ucb_beta=ucb_beta,
                                     state=model["state"])

    def _pairs(self, X, X2, want: Sequence[str]):
        fx = self.flatten(X)
        fx2 = fx if X2 is None else self.flatten(X2)
        state = sot.EvalState(fx2.trees, fx2.grammar)
        return sot.sot_pairs(fx.trees, state, self.device_weights(), float(self.variance), float(self.lengthscale), want=want)

    # ---- reference API -------------------------------------------------------------------
    def K(self, X, X2=None):
        """variance * exp(-d / lengthscale^2), d NOT squared (reference :110-113).  [len(X), len(X2)]."""
        return self._pairs(X, X2, ("K",))["K"]

    def K_diag(self, X):
        n = len(X)
        return torch.full((n,), float(self.variance), dtype=torch.float64, device=self._dev())

    def get_distance_matrix(self, X, X2=None):
        return self.wasserstein_distance(X, X2)

    def wasserstein_distance(self, X, X2=None):
        """sum_f alpha_f / sum(alpha) * L1_f (reference :124-162)."""
        return self._pairs(X, X2, ("D",))["D"]

    def per_feature_distances(self, X, X2=None):
        """[3, N, M] stack (DIM_WISE, PATHS, SUBTREES) of the per-feature Manhattan distances."""
        return self._pairs(X, X2, ("Df",))["Df"]

    def internal_transform_X(self, X, feature_type: FeatureType):
        """Sparse histograms of one feature type as a list of dicts key -> count / value, keys in the
        reference's insertion order (reference :164-185).  Keys are canonical ids, not Python hashes."""
        fx = self.flatten(X)
        f = {k: v.cpu().numpy() for k, v in sot.featurize(fx.trees, fx.grammar).items()}
        out = []
        for i in range(len(fx)):
            if feature_type == FeatureType.SUBTREES:
                d = {int(f["sub_id"][i, k]) & 0xFFFFFFFFFFFFFFFF: [int(f["sub_cnt"][i, k]), None] for k in range(int(f["sub_n"][i]))}
            elif feature_type == FeatureType.REDUCED_ELEMENTARY_PATHS:
                d = {int(f["path_id"][i, k]) & 0xFFFF: [int(f["path_cnt"][i, k]), None] for k in range(int(f["path_n"][i]))}
            else:
                d = {name: float(f["dimvec"][i, c]) for c, name in enumerate(fx.grammar.dim_col_names)}
            self.check_feature_dict(d)
            out.append(d)
        return out

    def check_feature_dict(self, feature_dict):
        assert any((v[0] if isinstance(v, list) else v) > 0 for v in feature_dict.values())

    def normalize_features(self, feature_type):
        return feature_type != FeatureType.DIM_WISE_WEIGHTED_ELEMENTARY_COUNT

    def create_hash_array_mapping(self, tree_dict_list):
        index_dict = {}
        for d in tree_dict_list:
            for key in d:
                if key not in index_dict:
                    index_dict[key] = [len(index_dict)]
        return index_dict

    def feature_dict_to_feature_vector(self, feature_dict, index_dict):
        v = np.zeros(len(index_dict))
        for key, val in feature_dict.items():
            v[index_dict[key][0]] = val[0] if isinstance(val, list) else val
        return v

    def feature_matrix(self, X, index_dict, normalize: bool) -> np.ndarray:
        rows = []
        for d in X:
            v = self.feature_dict_to_feature_vector(d, index_dict)
            rows.append(v / np.sum(v) if normalize else v)
        return np.array(rows)

    def dense_feature_matrices(self, X, X2=None):
        """All three dense [N, F] matrices at once, columns in the reference's first-seen order."""
        fx = self.flatten(X)
        f1 = sot.featurize(fx.trees, fx.grammar)
        f2 = None
        if X2 is not None:
            fx2 = self.flatten(X2)
            f2 = sot.featurize(fx2.trees, fx2.grammar)
        return sot.dense_feature_matrices(f1, f2)
