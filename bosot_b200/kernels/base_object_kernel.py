"""Kernel over arbitrary objects: the calling convention GPflow uses on the reference's
``BaseObjectKernel`` (bosot/kernels/base_object_kernel.py:23-44) without the gpflow base class -- there is
no GPflow on this path.

    kernel(X)                  -> K(X)            full Gram of one list
    kernel(X, X2)              -> K(X, X2)        cross Gram
    kernel(X, full_cov=False)  -> K_diag(X)       diagonal only
    kernel(X, X2, full_cov=False)                 ValueError, as in GPflow
"""
from abc import ABC, abstractmethod


class BaseObjectKernel(ABC):
    @abstractmethod
    def K(self, X, X2=None):
        """Gram matrix between the objects of X and of X2 (X itself when X2 is None)."""

    @abstractmethod
    def K_diag(self, X):
        """Diagonal of K(X)."""

    def transform_X(self, X):
        """Hook for kernels that canonicalise their inputs first; identity by default."""
        return X

    def __call__(self, X, X2=None, *, full_cov=True):
        if full_cov:
            return self.K(X, X2)
        if X2 is not None:
            raise ValueError("Ambiguous inputs: `not full_cov` and `X2` are not compatible.")
        return self.K_diag(X)
