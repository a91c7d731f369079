"""Mirror of bosot/configs/kernels/kernel_grammar_generators/compositional_kernel_search_configs.py."""
from .generator_configs import *  # noqa: F401,F403
