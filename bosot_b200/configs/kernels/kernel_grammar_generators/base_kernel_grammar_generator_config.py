from ...base import REQUIRED, BaseConfig


class BaseKernelGrammarGeneratorConfig(BaseConfig):
    input_dimension = REQUIRED
    name = REQUIRED
