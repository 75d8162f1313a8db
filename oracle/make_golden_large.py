"""Generate tests/golden/step_256.pt and step_512.pt by EXECUTING THE UNMODIFIED REFERENCE at the BASELINE sizes.

Run here, where /root/reference exists:   python oracle/make_golden_large.py
(the GPU box has no /root/reference; tests only read the committed fixtures).

One teacher-forced step at 256x256 (BASELINE configs[0]) and 512x512 (configs[1]), factor 4, driven exactly like
oracle/make_golden.py (closure = DIP.py:47-69).  The perturbed input z0 (33.5 MB at 512x512) is NOT stored: it is a
deterministic function of the seed --

    torch.manual_seed(seed); net = get_net(...); net_input = get_noise(32, 'noise', (H, W))
    z0 = net_input + net_input.clone().normal_() * 0.05

-- so the fixture keeps the seed and checksums of net_input / z0, and the tests rebuild z0 the same way (same torch
build, CPU generator).  Stored: out_HR, out_LR, loss, per-key gradient norms, full gradients of the small tensors,
[8, 8, k, k] slices of the conv gradients, and per-key projections of every gradient on a fixed pseudo-random
direction (a 112-number fingerprint of the whole 2.2 M-element gradient).
"""
import os
import sys

import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def checksum(t):
    t = t.detach().double().flatten()
    return (float(t.sum()), float(t.abs().sum()),
            float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64)).sum() / max(1, t.numel())))


def probe(k, shape):
    """Fixed pseudo-random direction for parameter `k` (shared with the tests)."""
    g = torch.Generator().manual_seed(abs(hash_str(k)) % (2 ** 31))
    return torch.randn(shape, generator=g, dtype=torch.float64)


def hash_str(s):
    h = 0
    for ch in s:
        h = (h * 131 + ord(ch)) % 2147483647
    return h


def main():
    sys.path.insert(0, REF)
    from models.DIP import get_net                      # noqa: E402
    from utils.downsampler import Downsampler           # noqa: E402
    from utils.DIP import get_noise, get_params, optimize  # noqa: E402

    torch.set_num_threads(1)   # fixes the reduction order of the fixtures
    for (H, seed) in ((256, 11), (512, 12)):
        W, factor, reg = H, 4, 0.05
        torch.manual_seed(seed)
        net = get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                      upsample_mode='bilinear')
        net_input = get_noise(32, 'noise', (H, W)).detach()
        net_input_saved = net_input.detach().clone()
        noise = net_input.detach().clone()
        # the synthetic image of the benchmark: low-frequency field + checker, reference downsampler
        g = torch.Generator().manual_seed(1000 + seed)
        low = torch.rand(1, 3, H // 8, H // 8, generator=g)
        field = torch.nn.functional.interpolate(low, size=(H, H), mode='bicubic', align_corners=False)[0]
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(H), indexing='ij')
        checker = (((yy // 16) + (xx // 16)) % 2).float() - 0.5
        hr = (field * 0.7 + 0.15 + 0.3 * checker).clamp(0, 1).unsqueeze(0)
        ds = Downsampler(n_planes=3, factor=factor, kernel_type='lanczos2', phase=0.5, preserve_size=True)
        with torch.no_grad():
            lr_img = ds(hr)
        msef = torch.nn.MSELoss()
        rec = {}

        def closure():
            z = net_input_saved + (noise.normal_() * reg)          # DIP.py:52
            rec['z0'] = z.clone()
            out_hr = net(z)                                         # DIP.py:60
            out_lr = ds(out_hr)                                     # DIP.py:62
            loss = msef(out_lr, lr_img)                             # DIP.py:65
            loss.backward()                                         # DIP.py:68
            rec['out_hr'] = out_hr.detach().clone()
            rec['out_lr'] = out_lr.detach().clone()
            rec['grads'] = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
            rec['loss'] = float(loss)
            return loss

        params = get_params('net', net, net_input)
        optimize('adam', params, closure, 0.01, 1)
        sd1 = {k: v.detach().clone() for k, v in net.state_dict().items()}
        grads = rec['grads']
        fixture = {
            'H': H, 'W': W, 'factor': factor, 'seed': seed, 'reg_noise_std': reg, 'lr': 0.01,
            'net_input_checksum': checksum(net_input_saved), 'z0_checksum': checksum(rec['z0']),
            'lr_img': lr_img, 'out_hr': rec['out_hr'], 'out_lr': rec['out_lr'], 'loss': rec['loss'],
            'grad_norms': {k: float(v.double().norm()) for k, v in grads.items()},
            'grad_full': {k: v for k, v in grads.items() if v.numel() <= 4096},
            'grad_slices': {k: v[:8, :8].clone() for k, v in grads.items() if v.dim() == 4 and v.shape[-1] == 3},
            'grad_probe': {k: float((v.double() * probe(k, v.shape)).sum()) for k, v in grads.items()},
            'post_adam_checksums': {k: checksum(v) for k, v in sd1.items() if v.dtype.is_floating_point},
            'post_adam_small': {k: v for k, v in sd1.items() if v.numel() <= 132},
        }
        path = os.path.join(OUT, f'step_{H}.pt')
        torch.save(fixture, path)
        print(f'step_{H}: loss {rec["loss"]:.6f}  {os.path.getsize(path)} bytes', flush=True)


if __name__ == '__main__':
    main()
