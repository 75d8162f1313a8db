"""Drop-in for models/GAN/discriminator.py: `from models.GAN.discriminator import Discriminator` (train_GAN.py:11)."""
from dsr_b200.gan_train import Discriminator, DiscriminatorConvBlock  # noqa: F401
