"""Name -> config class lookup (mirror of bosot/configs/config_picker.py:44-113) for the configs of the hot path:
the kernel-kernel configs, the GP-over-kernels model config, the BO configs and the search-space generator configs.
The reference's data-space model configs and greedy / TreeGEP search configs belong to parts that are out of scope
here and are not listed."""
from .bayesian_optimization import bayesian_optimizer_objects_configs as _bo
from .kernels.elementary_configs import (
    BasicLinearConfig, BasicPeriodicConfig, BasicRBFConfig, BasicRQConfig, LinearWithPriorConfig, PeriodicWithPriorConfig,
    RBFWithPriorConfig, RQWithPriorConfig)
from .kernels.grammar_tree_kernel_kernel_configs import OptimalTransportGrammarKernelConfig, OTWeightedDimsExtendedGrammarKernelConfig
from .kernels.hellinger_kernel_kernel_configs import BasicHellingerKernelKernelConfig
from .kernels.kernel_grammar_generators.generator_configs import (
    CKSHighDimGeneratorConfig, CKSWithRQGeneratorConfig, CompositionalKernelSearchGeneratorConfig, NDimFullKernelsGrammarGeneratorConfig)
from .models.object_gp_model_config import BasicObjectGPModelConfig


class ConfigPicker:
    models_configs_dict = {c.__name__: c for c in [BasicObjectGPModelConfig]}

    kernels_configs_dict = {
        c.__name__: c
        for c in [
            RBFWithPriorConfig, BasicRBFConfig, BasicLinearConfig, LinearWithPriorConfig, BasicPeriodicConfig, PeriodicWithPriorConfig,
            BasicRQConfig, RQWithPriorConfig, BasicHellingerKernelKernelConfig, OptimalTransportGrammarKernelConfig,
            OTWeightedDimsExtendedGrammarKernelConfig,
        ]
    }

    bayesian_optimization_configs_dict = {
        name: getattr(_bo, name)
        for name in ("ObjectBOExpectedImprovementConfig", "ObjectBOExpectedImprovementEAConfig",
                     "ObjectBOExpectedImprovementPerSecondEAConfig", "ObjectBOExpectedImprovementEAFewerStepsConfig", "ObjectBOGPUCBConfig")
        if hasattr(_bo, name)
    }

    kernel_grammar_generator_configs_dict = {
        c.__name__: c
        for c in [NDimFullKernelsGrammarGeneratorConfig, CompositionalKernelSearchGeneratorConfig, CKSWithRQGeneratorConfig, CKSHighDimGeneratorConfig]
    }

    @staticmethod
    def pick_kernel_config(config_class_name):
        return ConfigPicker.kernels_configs_dict[config_class_name]

    @staticmethod
    def pick_model_config(config_class_name):
        return ConfigPicker.models_configs_dict[config_class_name]

    @staticmethod
    def pick_bayesian_optimization_config(config_class_name):
        return ConfigPicker.bayesian_optimization_configs_dict[config_class_name]

    @staticmethod
    def pick_kernel_grammar_generator_config(config_class_name):
        return ConfigPicker.kernel_grammar_generator_configs_dict[config_class_name]
