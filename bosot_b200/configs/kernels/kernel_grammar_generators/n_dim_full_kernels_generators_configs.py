"""Generator config of the NDimFullKernelsSearchSpace: base kernels that act on all input dimensions at once
(bosot/configs/kernels/kernel_grammar_generators/n_dim_full_kernels_generators_configs.py:21-29)."""
from .base_kernel_grammar_generator_config import BaseKernelGrammarGeneratorConfig


class NDimFullKernelsGrammarGeneratorConfig(BaseKernelGrammarGeneratorConfig):
    n_initial_factor_trailing = 5
    n_exploration_trailing = 15
    exploration_p_geometric = 1.0 / 3.0
    n_exploitation_trailing = 50
    walk_length_exploitation_trailing = 5
    do_random_walk_exploitation_trailing = False
    name = "NDimFullKernelsGrammarGenerator"
