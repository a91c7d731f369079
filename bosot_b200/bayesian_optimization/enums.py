"""Enumerations of the BO layer.  Member names and values are the reference's
(bosot/bayesian_optimization/enums.py:20-38) because configs and scripts refer to them by name."""
from enum import Enum

_MEMBERS = {
    "AcquisitionFunctionType": ("GP_UCB", "EXPECTED_IMPROVEMENT", "RANDOM", "EXPECTED_IMPROVEMENT_PER_SECOND"),
    "ValidationType": ("MAX_OBSERVED",),
    "AcquisitionOptimizationObjectBOType": ("TRAILING_CANDIDATES", "EVOLUTIONARY"),
}


def _make(name):
    # values count from 1 in declaration order, as in the reference
    return Enum(name, {member: i + 1 for i, member in enumerate(_MEMBERS[name])}, module=__name__)


AcquisitionFunctionType = _make("AcquisitionFunctionType")
AcquisitionOptimizationObjectBOType = _make("AcquisitionOptimizationObjectBOType")
ValidationType = _make("ValidationType")

_VALIDATION_LABELS = {ValidationType.MAX_OBSERVED: "MAXOBSERVED"}
ValidationType.get_name = staticmethod(lambda val_type: _VALIDATION_LABELS.get(val_type))
