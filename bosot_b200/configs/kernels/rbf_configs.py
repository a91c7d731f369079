"""Mirror of bosot/configs/kernels/rbf_configs.py (names only; see elementary_configs.py)."""
from .elementary_configs import *  # noqa: F401,F403
