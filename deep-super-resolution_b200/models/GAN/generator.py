"""Drop-in for models/GAN/generator.py: `from models.GAN.generator import Generator` (eval_GAN.py:11)."""
from dsr_b200.gan import Generator, ResidualBlock, PixelShuffleBlock  # noqa: F401
