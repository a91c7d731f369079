"""bosot/configs/kernels/kernel_grammar_generators/*.py (all four share the same defaults)."""
from .base_kernel_grammar_generator_config import BaseKernelGrammarGeneratorConfig


class _TrailingDefaults(BaseKernelGrammarGeneratorConfig):
    n_initial_factor_trailing = 5
    n_exploration_trailing = 15
    exploration_p_geometric = 1.0 / 3.0
    n_exploitation_trailing = 50
    walk_length_exploitation_trailing = 5
    do_random_walk_exploitation_trailing = False


class CKSWithRQGeneratorConfig(_TrailingDefaults):
    name = "CKSWithRQGenerator"


class CKSHighDimGeneratorConfig(_TrailingDefaults):
    name = "CKSHighDimGenerator"


class CompositionalKernelSearchGeneratorConfig(_TrailingDefaults):
    name = "CompositionalKernelSearchGenerator"


class NDimFullKernelsGrammarGeneratorConfig(_TrailingDefaults):
    name = "NDimFullKernelsGrammarGenerator"
