"""Mirror of bosot/configs/kernels/kernel_grammar_generators/cks_with_rq_generator_config.py."""
from .generator_configs import *  # noqa: F401,F403
