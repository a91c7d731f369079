from ..base import REQUIRED, BaseConfig


class BaseKernelConfig(BaseConfig):
    """bosot/configs/kernels/base_kernel_config.py"""

    input_dimension = REQUIRED
    name = REQUIRED
