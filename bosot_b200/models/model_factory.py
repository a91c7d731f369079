"""bosot/models/model_factory.py:31-46 (object GP branch)."""
from ..configs.models.object_gp_model_config import BasicObjectGPModelConfig
from ..kernels.kernel_factory import KernelFactory
from .object_gp_model import ObjectGpModel


class ModelFactory:
    @staticmethod
    def build(model_config):
        if isinstance(model_config, BasicObjectGPModelConfig):
            kernel = KernelFactory.build(model_config.kernel_config)
            return ObjectGpModel(kernel=kernel, **model_config.dict())
        raise NotImplementedError(type(model_config).__name__)
