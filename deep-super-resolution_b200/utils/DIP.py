"""Drop-in for the reference's `utils.DIP` (utils/DIP.py): optimize, get_params, fill_noise, get_noise.
The reference module also re-exports utils.common via `from .common import *` (utils/DIP.py:3); that file
is host-side IO and stays the reference's own -- it is imported here when it is reachable."""
import numpy as np  # noqa: F401
import torch  # noqa: F401

try:  # utils/common.py of the reference checkout (namespace-package fall-through); optional
    from .common import *  # noqa: F401,F403
except Exception:  # pragma: no cover - reference checkout absent
    pass
from dsr_b200.optim import optimize, get_params, fill_noise, get_noise  # noqa: F401,E402
