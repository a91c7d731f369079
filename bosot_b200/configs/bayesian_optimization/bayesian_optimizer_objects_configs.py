"""bosot/configs/bayesian_optimization/bayesian_optimizer_objects_configs.py:21-57."""
from ...bayesian_optimization.enums import AcquisitionFunctionType, AcquisitionOptimizationObjectBOType, ValidationType
from ..base import REQUIRED, BaseConfig


class BaseObjectBOConfig(BaseConfig):
    acquisition_function_type = REQUIRED
    validation_type = ValidationType.MAX_OBSERVED
    acquisiton_optimization_type = AcquisitionOptimizationObjectBOType.TRAILING_CANDIDATES
    population_evolutionary = 100
    n_steps_evolutionary = 10
    num_offspring_evolutionary = 9
    n_prune_trailing = 600
    do_plotting = False
    name = "BaseObjectBO"


class ObjectBOExpectedImprovementConfig(BaseObjectBOConfig):
    acquisition_function_type = AcquisitionFunctionType.EXPECTED_IMPROVEMENT
    name = "ExpectedImprovementBO"


class ObjectBOExpectedImprovementEAConfig(ObjectBOExpectedImprovementConfig):
    acquisiton_optimization_type = AcquisitionOptimizationObjectBOType.EVOLUTIONARY
    population_evolutionary = 100
    n_steps_evolutionary = 10
    num_offspring_evolutionary = 4
    name = "ExpectedImprovementBOEA"


class ObjectBOExpectedImprovementEAFewerStepsConfig(ObjectBOExpectedImprovementEAConfig):
    n_steps_evolutionary = 6
    name = "ObjectBOExpectedImprovementEAFewerSteps"


class ObjectBOGPUCBConfig(BaseObjectBOConfig):
    acquisition_function_type = AcquisitionFunctionType.GP_UCB
    name = "GPUCBBO"
