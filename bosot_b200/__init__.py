"""bosot_b200 -- B200-native (sm_100a) implementation of BOSOT's kernel-kernel hot path.

SOT distance matrix between kernel-grammar expression trees -> Gram -> GP-over-kernels
posterior -> Expected Improvement, behind the reference's Python API
(``bosot_b200.kernels``, ``bosot_b200.models``, ``bosot_b200.bayesian_optimization`` mirror the
``bosot.*`` module paths of boschresearch/bosot).  The compute runs in hand-written CUDA
reached through the C-ABI of ``include/bosot_b200.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"
