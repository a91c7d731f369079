"""Mirror of bosot/configs/kernels/kernel_grammar_generators/n_dim_full_kernels_generators_configs.py."""
from .generator_configs import *  # noqa: F401,F403
