"""Configs of the elementary RBF kernel (bosot/configs/kernels/rbf_configs.py:22-33).  The SOT kernel-kernel only needs the
name; the Hellinger kernel-kernel samples the hyper-parameters from the Gamma priors given here (concentration, rate),
which only the ``*WithPrior`` variant carries."""
from .base_elementary_kernel_config import BaseElementaryKernelConfig


class BasicRBFConfig(BaseElementaryKernelConfig):
    base_lengthscale = 1.0
    base_variance = 1.0
    add_prior = False
    lengthscale_prior_parameters = (2.0, 2.0)
    variance_prior_parameters = (2.0, 3.0)
    name = "BasicRBF"


class RBFWithPriorConfig(BasicRBFConfig):
    add_prior = True
    name = "RBFWithPrior"
