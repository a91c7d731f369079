"""All elementary-kernel configs in one namespace (the reference keeps them in four modules:
bosot/configs/kernels/{rbf,linear,periodic,rational_quadratic}_configs.py, mirrored next to this file)."""
from .linear_configs import BasicLinearConfig, LinearWithPriorConfig  # noqa: F401
from .periodic_configs import BasicPeriodicConfig, PeriodicWithPriorConfig  # noqa: F401
from .rational_quadratic_configs import BasicRQConfig, RQWithPriorConfig  # noqa: F401
from .rbf_configs import BasicRBFConfig, RBFWithPriorConfig  # noqa: F401
