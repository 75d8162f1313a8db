"""Drop-in for the reference's `models.DIP` (models/DIP/__init__.py:8): `from models.DIP import get_net`.
Put deep-super-resolution_b200/ ahead of the reference checkout on sys.path; `models` and `utils` are
namespace packages in the reference (no __init__.py), so everything this package does not provide
(models.GAN, utils.common, ...) still resolves to the reference's files."""
from dsr_b200.net import get_net, SkipNet  # noqa: F401
