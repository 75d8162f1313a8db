"""PSNR / SSIM of the logging branch (DIP.py:71-87): the CPU restatement of the torchmetrics algorithm
(oracle/metrics_oracle.py) on known answers, and the CUDA kernels (dsr_psnr / dsr_ssim through the C ABI) against it."""
import math

import pytest
import torch

from oracle import metrics_oracle as M


def test_metric_oracle_known_answers():
    g = torch.Generator().manual_seed(0)
    a = torch.rand(2, 3, 40, 56, generator=g)
    assert M.ssim(a, a) == pytest.approx(1.0, abs=1e-12)
    assert M.psnr(a + 0.1, a, data_range=1.0) == pytest.approx(20.0, abs=1e-4)
    # data_range=None: max(target) - min(min(target), 0)
    t = a * 0.5 + 0.2
    want = 10 * math.log10(float(t.max()) ** 2 / float(((t + 0.05 - t) ** 2).mean()))
    assert M.psnr(t + 0.05, t) == pytest.approx(want, abs=1e-3)
    w = M.gaussian_1d()
    assert float(w.sum()) == pytest.approx(1.0) and w.argmax() == 5 and torch.allclose(w, w.flip(0))
    # a constant offset lowers luminance similarity only: SSIM = (2 mu_a mu_b + c1) / (mu_a^2 + mu_b^2 + c1) for flat images
    f1, f2 = torch.full((1, 1, 32, 32), 0.3), torch.full((1, 1, 32, 32), 0.5)
    assert M.ssim(f1, f2) == pytest.approx((2 * 0.15 + 1e-4) / (0.09 + 0.25 + 1e-4), rel=1e-6)
    b = (a + 0.2 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    assert 0.0 < M.ssim(b, a) < 0.9


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(1, 3, 512, 512), (2, 3, 72, 88), (1, 3, 33, 47)])
def test_cuda_metrics_match_the_restatement(shape):
    import dsr_b200
    g = torch.Generator().manual_seed(1)
    t = torch.rand(shape, generator=g) * 0.9
    p = (t + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
    psnr = dsr_b200.PeakSignalNoiseRatio().to('cuda')
    ssim = dsr_b200.StructuralSimilarityIndexMeasure(data_range=1.).to('cuda')
    for _ in range(2):                                   # the workspace re-arms itself: the second call must agree
        assert psnr(p.cuda(), t.cuda()).item() == pytest.approx(M.psnr(p, t), abs=2e-4)
        assert ssim(p.cuda(), t.cuda()).item() == pytest.approx(M.ssim(p, t), abs=2e-5)
    assert dsr_b200.PeakSignalNoiseRatio(data_range=1.0)(p.cuda(), t.cuda()).item() == pytest.approx(M.psnr(p, t, 1.0), abs=2e-4)
    assert ssim(t.cuda(), t.cuda()).item() == pytest.approx(1.0, abs=1e-6)
    with pytest.raises(RuntimeError):
        psnr(p, t)                                       # CPU tensors: no fallback


@pytest.mark.gpu
def test_torchmetrics_drop_in_names():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'deep-super-resolution_b200')
    code = ('import sys; sys.path[:0] = [%r, %r]\n'
            'from torchmetrics.image import PeakSignalNoiseRatio as PSNR, StructuralSimilarityIndexMeasure as SSIM\n'
            'from torchmetrics.image.lpip import LearnedPerceptualImagePatchSimilarity as LPIPS\n'
            'import torch, dsr_b200\n'
            'assert PSNR is dsr_b200.PeakSignalNoiseRatio and SSIM is dsr_b200.StructuralSimilarityIndexMeasure\n'
            'a = torch.rand(1, 3, 64, 64, device="cuda"); print(SSIM(data_range=1.).to("cuda")(a, a).item())\n'
            % (os.path.join(pkg, 'metrics_dropin'), pkg))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, check=True).stdout
    assert float(out.strip().splitlines()[-1]) == pytest.approx(1.0, abs=1e-6)
