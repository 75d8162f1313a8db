"""Drop-in for the reference's `utils.DIP` (utils/DIP.py): optimize, get_params, fill_noise, get_noise.
The reference module also re-exports utils.common via `from .common import *` (utils/DIP.py:3); so does this one (the
drop-in `utils/common.py` beside it)."""
import numpy as np  # noqa: F401
import torch  # noqa: F401

from .common import *  # noqa: F401,F403
from dsr_b200.optim import optimize, get_params, fill_noise, get_noise  # noqa: F401,E402
