"""Generator config of the CompositionalKernelSearchSpace: RBF, periodic and linear
(bosot/configs/kernels/kernel_grammar_generators/compositional_kernel_search_configs.py:20-28)."""
from .base_kernel_grammar_generator_config import BaseKernelGrammarGeneratorConfig


class CompositionalKernelSearchGeneratorConfig(BaseKernelGrammarGeneratorConfig):
    n_initial_factor_trailing = 5
    n_exploration_trailing = 15
    exploration_p_geometric = 1.0 / 3.0
    n_exploitation_trailing = 50
    walk_length_exploitation_trailing = 5
    do_random_walk_exploitation_trailing = False
    name = "CompositionalKernelSearchGenerator"
