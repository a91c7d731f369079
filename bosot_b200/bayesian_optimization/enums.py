"""bosot/bayesian_optimization/enums.py:20-38."""
from enum import Enum


class AcquisitionFunctionType(Enum):
    GP_UCB = 1
    EXPECTED_IMPROVEMENT = 2
    RANDOM = 3
    EXPECTED_IMPROVEMENT_PER_SECOND = 4


class ValidationType(Enum):
    MAX_OBSERVED = 1

    @staticmethod
    def get_name(val_type):
        if val_type == ValidationType.MAX_OBSERVED:
            return "MAXOBSERVED"


class AcquisitionOptimizationObjectBOType(Enum):
    TRAILING_CANDIDATES = 1
    EVOLUTIONARY = 2
