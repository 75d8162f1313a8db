"""dsr_b200 -- B200-native Deep-Image-Prior super-resolution step.

Python mirror of the reference call surface (LewisClifton/Deep-Super-Resolution: models/DIP,
utils/downsampler.py, utils/DIP.py and the closure of DIP.py) on top of libdsr_b200.so, a
hand-written sm_100a CUDA library reached through a plain C ABI (include/dsr_b200.h).
"""
from . import _lib                                   # noqa: F401  (fails loudly when the library is missing)
from .net import SkipNet, get_net                    # noqa: F401
from .downsampler import Downsampler, get_kernel     # noqa: F401
from .optim import optimize, get_params, get_noise, fill_noise   # noqa: F401
from .dip import DIP_ISR, dip_sr_fused                # noqa: F401
from .gan import Generator                            # noqa: F401
from .metrics import PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure   # noqa: F401
from .evalgan import to_uint8_hwc, GAN_ISR_Batch_eval   # noqa: F401
