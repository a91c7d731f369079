"""Configs of the elementary rational-quadratic kernel (bosot/configs/kernels/rational_quadratic_configs.py:22-35).  The SOT kernel-kernel only needs the
name; the Hellinger kernel-kernel samples the hyper-parameters from the Gamma priors given here (concentration, rate),
which only the ``*WithPrior`` variant carries."""
from .base_elementary_kernel_config import BaseElementaryKernelConfig


class BasicRQConfig(BaseElementaryKernelConfig):
    base_lengthscale = 1.0
    base_variance = 1.0
    base_alpha = 1.0
    add_prior = False
    lengthscale_prior_parameters = (2.0, 2.0)
    variance_prior_parameters = (2.0, 3.0)
    alpha_prior_parameters = (2.0, 2.0)
    name = "BasicRQ"


class RQWithPriorConfig(BasicRQConfig):
    add_prior = True
    name = "RQWithPrior"
