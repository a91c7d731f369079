"""Mirror of bosot/configs/kernels/periodic_configs.py (names only; see elementary_configs.py)."""
from .elementary_configs import *  # noqa: F401,F403
