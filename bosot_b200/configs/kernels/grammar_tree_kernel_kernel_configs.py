"""bosot/configs/kernels/grammar_tree_kernel_kernel_configs.py:26-41."""
from ...kernels.kernel_kernel_grammar_tree import FeatureType
from .base_kernel_config import BaseKernelConfig
from ..base import REQUIRED


class OptimalTransportGrammarKernelConfig(BaseKernelConfig):
    feature_type_list = REQUIRED
    input_dimension = 0
    base_variance = 1.0
    base_lengthscale = 1.0
    base_alpha = 0.5
    alpha_trainable = True
    parameters_trainable = True
    transform_to_normal = False
    name = "OptimalTransportGrammarKernel"


class OTWeightedDimsExtendedGrammarKernelConfig(OptimalTransportGrammarKernelConfig):
    feature_type_list = [
        FeatureType.DIM_WISE_WEIGHTED_ELEMENTARY_COUNT,
        FeatureType.REDUCED_ELEMENTARY_PATHS,
        FeatureType.SUBTREES,
    ]
    name = "OTWeightedDimsExtendedGrammarKernel"
