"""The evaluation loop of eval_GAN.py:21-67 (`GAN_ISR_Batch_eval`) over the B200 library: generator inference
(`dsr_b200.gan.Generator`, eval mode), on-device PSNR / SSIM, and the image writer path -- the resolved image becomes the
H x W x C uint8 array of eval_GAN.py:50-53 ON THE DEVICE (`to_uint8_hwc`, one launch of csrc/dsr_metrics.cu), so a quarter
of the bytes crosses PCIe and the host only encodes the PNG (`utils.common.save_image`, utils/common.py:20-33)."""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch

from . import _lib
from ._lib import lib, check
from .metrics import PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure


def to_uint8_hwc(image: torch.Tensor, clip: bool = False) -> torch.Tensor:
    """[B, C, H, W] (or [C, H, W]) fp32 on a CUDA device -> [B, H, W, C] (or [H, W, C]) uint8 on the same device.

    clip=False: ``(x.transpose(1, 2, 0) * 255).astype(np.uint8)`` of eval_GAN.py:52, bit for bit (numpy truncates towards
    zero and keeps the low byte, so the negative values of a tanh output wrap); clip=True: ``np.clip(x * 255, 0, 255)
    .astype(np.uint8)`` of utils/common.py:81 (np_to_pil)."""
    _lib.require_cuda()
    if not image.is_cuda:
        raise RuntimeError('dsr_b200.to_uint8_hwc runs on a CUDA device only (no CPU fallback)')
    squeeze = image.dim() == 3
    x = image.detach().float()
    x = (x.unsqueeze(0) if squeeze else x).contiguous()
    if x.dim() != 4:
        raise ValueError(f'expected [B, C, H, W] or [C, H, W], got {tuple(image.shape)}')
    B, Cn, H, W = x.shape
    out = torch.empty((B, H, W, Cn), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.dsr_image_to_u8_hwc(x.data_ptr(), B, Cn, H, W, int(bool(clip)), out.data_ptr(), _lib.stream_ptr()),
              'dsr_image_to_u8_hwc')
    return out[0] if squeeze else out


def GAN_ISR_Batch_eval(gan_G, val_loader, out_dir, batch_size, device, lpips: Optional[Callable] = None,
                       save_image: Optional[Callable] = None):
    """eval_GAN.py:21-67 with the same positional arguments and the same result dictionary.  ``lpips`` (optional
    callable) stands for the pretrained-AlexNet LPIPS object of eval_GAN.py:32 (third party, out of scope): without it
    ``avg_lpips`` is None.  ``save_image`` defaults to the drop-in ``utils.common.save_image``."""
    if save_image is None:
        from utils.common import save_image                  # the drop-in (or the reference's own) writer
    psnr = PeakSignalNoiseRatio().to(device)                 # eval_GAN.py:30-31
    ssim = StructuralSimilarityIndexMeasure(data_range=1.).to(device)
    running_psnr = running_ssim = 0.0
    running_lpips = 0.0 if lpips is not None else None
    for LR_image, HR_image, image_name in val_loader:
        HR_image = HR_image.to(device)
        LR_image = LR_image.to(device)
        image_name = image_name[0]
        print(f'Starting on {image_name}.')
        with torch.no_grad():
            resolved_image = gan_G(LR_image)
        m_psnr, m_ssim = psnr(resolved_image, HR_image), ssim(resolved_image, HR_image)
        if lpips is not None:
            running_lpips += float(lpips(resolved_image, HR_image))
        pixels = to_uint8_hwc(resolved_image[0])              # eval_GAN.py:50-52 on the device
        running_psnr += m_psnr.item()
        running_ssim += m_ssim.item()
        print(f'Done evaluating over {image_name}.')
        save_image(pixels.cpu().numpy(), image_name, out_dir)
        print()
    return {'avg_psnr': running_psnr / batch_size, 'avg_ssim': running_ssim / batch_size,
            'avg_lpips': (running_lpips / batch_size) if running_lpips is not None else None}
