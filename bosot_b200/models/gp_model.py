"""bosot/models/gp_model.py:37-39 (only what the object GP needs from it)."""
from enum import Enum


class PredictionQuantity(Enum):
    PREDICT_F = 0
    PREDICT_Y = 1
