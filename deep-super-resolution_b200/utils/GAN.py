"""Drop-in for the reference's `utils.GAN` (utils/GAN.py): Vgg19Loss, PerceptualLoss, get_loss_D, get_adversarial_loss.
train_GAN.py:14 does `from utils.GAN import *` and relies on the names `nn`, `torch`, `os` it re-exports."""
import os  # noqa: F401

import torch  # noqa: F401
import torch.nn as nn  # noqa: F401

from dsr_b200.gan_train import Vgg19Loss, PerceptualLoss, get_loss_D, get_adversarial_loss  # noqa: F401
