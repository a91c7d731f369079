"""bosot/configs/kernels/hellinger_kernel_kernel_configs.py:21-30."""
from .base_kernel_config import BaseKernelConfig


class BasicHellingerKernelKernelConfig(BaseKernelConfig):
    base_lengthscale = 0.5
    base_variance = 1.0
    num_param_samples = 100
    num_virtual_points = 20
    use_sobol_virtual_points = False
    use_hyperprior = False
    lengthscale_prior_parameters = (0.1, 0.7)
    variance_prior_parameters = (0.4, 0.7)
    name = "BasicHellingerKernelKernel"
