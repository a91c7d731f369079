"""bosot/kernels/kernel_grammar/generator_factory.py:31-49."""
from ...configs.kernels.kernel_grammar_generators.generator_configs import (
    CKSHighDimGeneratorConfig,
    CKSWithRQGeneratorConfig,
    CompositionalKernelSearchGeneratorConfig,
    NDimFullKernelsGrammarGeneratorConfig,
)
from .kernel_grammar_candidate_generator import KernelGrammarCandidateGenerator
from .kernel_grammar_search_spaces import CKSHighDimSearchSpace, CKSWithRQSearchSpace, CompositionalKernelSearchSpace, NDimFullKernelsSearchSpace

_SPACES = [
    (NDimFullKernelsGrammarGeneratorConfig, NDimFullKernelsSearchSpace),
    (CompositionalKernelSearchGeneratorConfig, CompositionalKernelSearchSpace),
    (CKSWithRQGeneratorConfig, CKSWithRQSearchSpace),
    (CKSHighDimGeneratorConfig, CKSHighDimSearchSpace),
]


class GeneratorFactory:
    @staticmethod
    def build(generator_config):
        for cfg_cls, space_cls in _SPACES:
            if isinstance(generator_config, cfg_cls):
                return KernelGrammarCandidateGenerator(search_space=space_cls(generator_config.input_dimension), **generator_config.dict())
        raise NotImplementedError(type(generator_config).__name__)
