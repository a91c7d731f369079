"""Mirror of bosot/configs/kernels/rational_quadratic_configs.py (names only; see elementary_configs.py)."""
from .elementary_configs import *  # noqa: F401,F403
