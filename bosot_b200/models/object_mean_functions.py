"""Mean functions over objects (mirror of bosot/models/object_mean_functions.py:29-48)."""
import numpy as np

from ..kernels.kernel_kernel_grammar_tree import Parameter


class MeanFunction:
    def value(self) -> float:
        raise NotImplementedError

    def __call__(self, X):
        return np.full((len(X), 1), self.value(), dtype=np.float64)


class Zero(MeanFunction):
    def __init__(self, output_dim=1):
        self.output_dim = output_dim

    def value(self) -> float:
        return 0.0


class ObjectConstant(MeanFunction):
    def __init__(self, base_c=0.0, trainable=True, output_dim=1):
        self.output_dim = output_dim
        self.c = Parameter(base_c, trainable=trainable)

    def value(self) -> float:
        return float(self.c)
