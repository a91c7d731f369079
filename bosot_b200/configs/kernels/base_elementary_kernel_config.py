from .base_kernel_config import BaseKernelConfig


class BaseElementaryKernelConfig(BaseKernelConfig):
    """bosot/configs/kernels/base_elementary_kernel_config.py"""

    active_on_single_dimension = False
    active_dimension = 0
