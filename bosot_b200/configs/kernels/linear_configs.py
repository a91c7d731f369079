"""Configs of the elementary linear kernel (bosot/configs/kernels/linear_configs.py:22-33).  The SOT kernel-kernel only needs the
name; the Hellinger kernel-kernel samples the hyper-parameters from the Gamma priors given here (concentration, rate),
which only the ``*WithPrior`` variant carries."""
from .base_elementary_kernel_config import BaseElementaryKernelConfig


class BasicLinearConfig(BaseElementaryKernelConfig):
    base_variance = 1.0
    base_offset = 1.0
    add_prior = False
    variance_prior_parameters = (2.0, 3.0)
    offset_prior_parameters = (2.0, 3.0)
    name = "BasicLinear"


class LinearWithPriorConfig(BasicLinearConfig):
    add_prior = True
    name = "LinearWithPrior"
