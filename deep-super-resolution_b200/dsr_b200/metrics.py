"""On-device PSNR / SSIM for the logging branch of the DIP loop (DIP.py:71-87, :183-185).

`PeakSignalNoiseRatio` and `StructuralSimilarityIndexMeasure` take the constructor arguments DIP.py:157-158 and
train_GAN.py:31-32 pass to the torchmetrics classes of the same names and are called the same way --
``psnr(out_HR, HR_image).item()`` -- but each call is ONE kernel launch of libdsr_b200.so (csrc/dsr_metrics.cu) that
leaves a 0-dim tensor on the device.  torchmetrics itself is a third-party dependency of the reference (absent from
this image); the tests compare against a CPU restatement of the algorithm it publishes.
`deep-super-resolution_b200/metrics_dropin/` exposes these classes under the module names DIP.py imports.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check


class _Metric(nn.Module):
    def __init__(self):
        super().__init__()
        self._ws = {}

    def _workspace(self, device: torch.device) -> torch.Tensor:
        ws = self._ws.get(device)
        if ws is None:
            ws = torch.zeros(max(64, lib.dsr_metric_workspace_bytes()), dtype=torch.uint8, device=device)
            self._ws[device] = ws
        return ws

    @staticmethod
    def _pair(preds: torch.Tensor, target: torch.Tensor):
        if not (preds.is_cuda and target.is_cuda):
            raise RuntimeError('dsr_b200 metrics run on a CUDA device only (no CPU fallback)')
        if preds.shape != target.shape:
            raise ValueError(f'preds {tuple(preds.shape)} and target {tuple(target.shape)} must have the same shape')
        return preds.detach().float().contiguous(), target.detach().float().contiguous()


class PeakSignalNoiseRatio(_Metric):
    """torchmetrics.image.PeakSignalNoiseRatio(data_range=None, base=10), value of one update (what forward returns)."""

    def __init__(self, data_range=None, base: float = 10.0, **kwargs):
        super().__init__()
        if base != 10.0 or kwargs.get('dim') is not None:
            raise NotImplementedError('dsr_b200 PeakSignalNoiseRatio supports base=10 over the whole tensor (DIP.py:157)')
        self.data_range = -1.0 if data_range is None else float(data_range)

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        p, t = self._pair(preds, target)
        out = torch.empty((), dtype=torch.float32, device=p.device)
        check(lib.dsr_psnr(p.data_ptr(), t.data_ptr(), p.numel(), self.data_range, self._workspace(p.device).data_ptr(),
                           out.data_ptr(), _lib.stream_ptr()), 'dsr_psnr')
        return out


class StructuralSimilarityIndexMeasure(_Metric):
    """torchmetrics.image.StructuralSimilarityIndexMeasure(gaussian_kernel=True, sigma=1.5, kernel_size=11,
    data_range=..., k1=0.01, k2=0.03), mean over the batch."""

    def __init__(self, data_range=None, gaussian_kernel: bool = True, sigma: float = 1.5, kernel_size: int = 11,
                 k1: float = 0.01, k2: float = 0.03, **kwargs):
        super().__init__()
        if not gaussian_kernel or sigma != 1.5 or kernel_size != 11 or k1 != 0.01 or k2 != 0.03:
            raise NotImplementedError('dsr_b200 StructuralSimilarityIndexMeasure supports the torchmetrics defaults '
                                      '(11x11 Gaussian window, sigma 1.5, k1 0.01, k2 0.03)')
        self.data_range = data_range

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        p, t = self._pair(preds, target)
        if p.dim() != 4:
            raise ValueError('expected [N, C, H, W] images')
        dr = self.data_range
        if dr is None:        # torchmetrics: max(preds.max() - preds.min(), target.max() - target.min())
            dr = float(torch.maximum(p.max() - p.min(), t.max() - t.min()))
        n, c, h, w = p.shape
        out = torch.empty((), dtype=torch.float32, device=p.device)
        check(lib.dsr_ssim(p.data_ptr(), t.data_ptr(), n * c, h, w, float(dr), self._workspace(p.device).data_ptr(),
                           out.data_ptr(), _lib.stream_ptr()), 'dsr_ssim')
        return out
