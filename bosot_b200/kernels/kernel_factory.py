"""config -> object factory (mirror of bosot/kernels/kernel_factory.py:34-54).

Elementary kernels are symbolic on this path: only their name (``<ConfigName>_on_<dim>``,
bosot/kernels/base_elementary_kernel.py:34-35) reaches the kernel-kernel; data-space GP kernels
belong to the oracle side, which is out of scope."""
from ..configs.kernels.base_elementary_kernel_config import BaseElementaryKernelConfig
from ..configs.kernels.grammar_tree_kernel_kernel_configs import OptimalTransportGrammarKernelConfig
from ..configs.kernels.hellinger_kernel_kernel_configs import BasicHellingerKernelKernelConfig
from .kernel_grammar.kernel_grammar import SymbolicKernel


class KernelFactory:
    @staticmethod
    def build(kernel_config):
        if isinstance(kernel_config, BaseElementaryKernelConfig):
            name = kernel_config.name
            if kernel_config.active_on_single_dimension:
                name = name + "_on_" + str(kernel_config.active_dimension)
            return SymbolicKernel(name, kernel_config.input_dimension)
        if isinstance(kernel_config, OptimalTransportGrammarKernelConfig):
            from .kernel_kernel_grammar_tree import OptimalTransportKernelKernel

            return OptimalTransportKernelKernel(**kernel_config.dict())
        if isinstance(kernel_config, BasicHellingerKernelKernelConfig):
            from .kernel_kernel_hellinger import KernelKernelHellinger

            return KernelKernelHellinger(**kernel_config.dict())
        raise NotImplementedError(type(kernel_config).__name__)
