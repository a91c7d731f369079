"""Abstract kernel over objects (mirror of bosot/kernels/base_object_kernel.py:23-44, minus the
gpflow base class: there is no GPflow on this path)."""
from abc import ABC, abstractmethod
from typing import List, Optional


class BaseObjectKernel(ABC):
    @abstractmethod
    def K(self, X: List[object], X2: Optional[List[object]] = None):
        raise NotImplementedError

    @abstractmethod
    def K_diag(self, X: List[object]):
        raise NotImplementedError

    def transform_X(self, X: List[object]):
        return X

    def __call__(self, X, X2=None, *, full_cov=True):
        if (not full_cov) and (X2 is not None):
            raise ValueError("Ambiguous inputs: `not full_cov` and `X2` are not compatible.")
        if not full_cov:
            return self.K_diag(X)
        return self.K(X, X2)
