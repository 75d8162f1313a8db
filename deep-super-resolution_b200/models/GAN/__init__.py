"""Drop-in for the reference's `models.GAN` package: provides `models.GAN.generator` (B200 inference path); every
other sub-module (models.GAN.discriminator, ...) still resolves to the reference checkout further down sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
