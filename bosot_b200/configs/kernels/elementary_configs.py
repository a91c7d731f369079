"""Elementary kernels (bosot/configs/kernels/{rbf,linear,periodic,rational_quadratic}_configs.py).  The SOT
kernel-kernel only needs their names; the Hellinger kernel-kernel samples their hyper-parameters from the Gamma
priors given here (concentration, rate), which only the ``*WithPrior`` variants carry."""
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


class BasicPeriodicConfig(BaseElementaryKernelConfig):
    base_lengthscale = 1.0
    base_variance = 1.0
    base_period = 1.0
    add_prior = False
    lengthscale_prior_parameters = (2.0, 2.0)
    variance_prior_parameters = (2.0, 3.0)
    period_prior_parameters = (2.0, 2.0)
    name = "BasicPeriodicKernel"


class PeriodicWithPriorConfig(BasicPeriodicConfig):
    add_prior = True
    name = "PeriodicKernelWithPrior"


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
