"""The four generator configs in one namespace (one module each in the reference, mirrored next to this file)."""
from .cks_high_dim_generator_config import CKSHighDimGeneratorConfig  # noqa: F401
from .cks_with_rq_generator_config import CKSWithRQGeneratorConfig  # noqa: F401
from .compositional_kernel_search_configs import CompositionalKernelSearchGeneratorConfig  # noqa: F401
from .n_dim_full_kernels_generators_configs import NDimFullKernelsGrammarGeneratorConfig  # noqa: F401
