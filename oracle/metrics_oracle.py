"""CPU restatement of the two torchmetrics image metrics DIP.py logs.  TEST INFRASTRUCTURE ONLY.

torchmetrics is a third-party dependency of the reference (DIP.py:7-8; unpinned: the reference has no requirements
file) and is NOT installed in this image, so these functions restate the algorithm torchmetrics 1.x publishes
(functional/image/psnr.py: _psnr_update / _psnr_compute; functional/image/ssim.py: _ssim_update with
gaussian_kernel=True, sigma=1.5, kernel_size=11, k1=0.01, k2=0.03).  PARITY UNPINNED against torchmetrics itself;
anchored on the reference's call sites DIP.py:71-79,157-158,183-185.
"""
import math

import torch
import torch.nn.functional as F


def psnr(preds: torch.Tensor, target: torch.Tensor, data_range=None) -> float:
    """PeakSignalNoiseRatio() as DIP.py:157 builds it, value of a single update: with data_range=None the range is
    max(target) - min(target, 0) (min_target / max_target states start at 0)."""
    if data_range is None:
        data_range = max(float(target.max()), 0.0) - min(float(target.min()), 0.0)
    mse = float(((preds.double() - target.double()) ** 2).mean())
    return 10.0 * (2.0 * math.log10(data_range) - math.log10(mse))


def gaussian_1d(kernel_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=torch.float64)
    g = torch.exp(-((dist / sigma) ** 2) / 2)
    return g / g.sum()


def ssim(preds: torch.Tensor, target: torch.Tensor, data_range: float = 1.0) -> float:
    """StructuralSimilarityIndexMeasure(data_range=...) with the torchmetrics defaults: reflect-pad by 5, 11 x 11
    Gaussian window as a grouped conv over (p, t, pp, tt, pt), variances clamped at 0, crop the 5-pixel border again,
    mean per image, mean over the batch."""
    p, t = preds.double(), target.double()
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    ch = p.shape[1]
    g = gaussian_1d()
    k2d = torch.outer(g, g).expand(ch, 1, 11, 11)
    pad = 5
    pp = F.pad(p, (pad, pad, pad, pad), mode='reflect')
    tp = F.pad(t, (pad, pad, pad, pad), mode='reflect')
    stack = torch.cat((pp, tp, pp * pp, tp * tp, pp * tp))
    out = F.conv2d(stack, k2d, groups=ch)
    n = p.shape[0]
    mu_p, mu_t, e_pp, e_tt, e_pt = (out[i * n:(i + 1) * n] for i in range(5))
    var_p = torch.clamp(e_pp - mu_p ** 2, min=0.0)
    var_t = torch.clamp(e_tt - mu_t ** 2, min=0.0)
    cov = e_pt - mu_p * mu_t
    full = ((2 * mu_p * mu_t + c1) * (2 * cov + c2)) / ((mu_p ** 2 + mu_t ** 2 + c1) * (var_p + var_t + c2))
    idx = full[..., pad:-pad, pad:-pad]
    return float(idx.reshape(n, -1).mean(-1).mean())
