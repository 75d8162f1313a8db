"""Drop-in for the reference's `utils.downsampler` (utils/downsampler.py): Downsampler, get_kernel."""
from dsr_b200.downsampler import Downsampler, get_kernel  # noqa: F401
