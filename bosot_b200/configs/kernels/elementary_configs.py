"""Names of the elementary kernels (the only thing the kernel-kernel path needs from them):
bosot/configs/kernels/{rbf,linear,periodic,rational_quadratic}_configs.py."""
from .base_elementary_kernel_config import BaseElementaryKernelConfig


class BasicRBFConfig(BaseElementaryKernelConfig):
    name = "BasicRBF"


class RBFWithPriorConfig(BasicRBFConfig):
    name = "RBFWithPrior"


class BasicLinearConfig(BaseElementaryKernelConfig):
    name = "BasicLinear"


class LinearWithPriorConfig(BasicLinearConfig):
    name = "LinearWithPrior"


class BasicPeriodicConfig(BaseElementaryKernelConfig):
    name = "BasicPeriodicKernel"


class PeriodicWithPriorConfig(BasicPeriodicConfig):
    name = "PeriodicKernelWithPrior"


class BasicRQConfig(BaseElementaryKernelConfig):
    name = "BasicRQ"


class RQWithPriorConfig(BasicRQConfig):
    name = "RQWithPrior"
