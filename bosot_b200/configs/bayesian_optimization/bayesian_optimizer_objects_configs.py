"""Presets of the object-BO loop.  Field names, defaults and class names are the reference's
(bosot/configs/bayesian_optimization/bayesian_optimizer_objects_configs.py:21-57): scripts build
``BayesianOptimizerObjects(**SomeConfig(...).dict())`` from them."""
from ...bayesian_optimization.enums import AcquisitionFunctionType as _Acq
from ...bayesian_optimization.enums import AcquisitionOptimizationObjectBOType as _Opt
from ...bayesian_optimization.enums import ValidationType as _Val
from ..base import REQUIRED, BaseConfig


def _preset(class_name, parent, **fields):
    """A config class = its parent plus overridden defaults."""
    return type(class_name, (parent,), dict(fields, __module__=__name__))


BaseObjectBOConfig = _preset(
    "BaseObjectBOConfig", BaseConfig,
    acquisition_function_type=REQUIRED,
    validation_type=_Val.MAX_OBSERVED,
    acquisiton_optimization_type=_Opt.TRAILING_CANDIDATES,
    population_evolutionary=100, n_steps_evolutionary=10, num_offspring_evolutionary=9,  # population % (offspring + 1) == 0
    n_prune_trailing=600, do_plotting=False, name="BaseObjectBO",
)
ObjectBOExpectedImprovementConfig = _preset(
    "ObjectBOExpectedImprovementConfig", BaseObjectBOConfig,
    acquisition_function_type=_Acq.EXPECTED_IMPROVEMENT, name="ExpectedImprovementBO")
ObjectBOExpectedImprovementEAConfig = _preset(
    "ObjectBOExpectedImprovementEAConfig", ObjectBOExpectedImprovementConfig,
    acquisiton_optimization_type=_Opt.EVOLUTIONARY, population_evolutionary=100, n_steps_evolutionary=10,
    num_offspring_evolutionary=4, name="ExpectedImprovementBOEA")
ObjectBOExpectedImprovementEAFewerStepsConfig = _preset(
    "ObjectBOExpectedImprovementEAFewerStepsConfig", ObjectBOExpectedImprovementEAConfig,
    n_steps_evolutionary=6, name="ObjectBOExpectedImprovementEAFewerSteps")
ObjectBOGPUCBConfig = _preset("ObjectBOGPUCBConfig", BaseObjectBOConfig, acquisition_function_type=_Acq.GP_UCB, name="GPUCBBO")
