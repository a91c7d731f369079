"""bosot/configs/models/object_gp_model_config.py:22-33."""
from ...models.gp_model import PredictionQuantity
from ..base import REQUIRED, BaseConfig


class BaseModelConfig(BaseConfig):
    name = REQUIRED


class BasicObjectGPModelConfig(BaseModelConfig):
    kernel_config = REQUIRED
    observation_noise = 0.01
    optimize_hps = True
    train_likelihood_variance = True
    pertube_parameters_at_start = True
    perform_multi_start_optimization = True
    n_starts_for_multistart_opt = 5
    set_prior_on_observation_noise = False
    expected_observation_noise = 0.1
    prediction_quantity = PredictionQuantity.PREDICT_Y
    name = "ObjectGPModel"
