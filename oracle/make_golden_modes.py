"""Teacher-forced step fixtures for the other two get_net options reachable through its signature:
pad='zero' (models/DIP/utils.py:96-102) and upsample_mode='nearest' (models/DIP/skip.py:77).  Build container only:
imports the UNMODIFIED reference from /root/reference.

    python oracle/make_golden_modes.py          -> tests/golden/step_<pad>_<upsample>_<H>x<W>.pt

Same recipe as oracle/make_golden.py (single-threaded for a fixed reduction order): torch.manual_seed(seed) -> get_net
-> get_noise; one DIP.py:47-69 closure evaluation with a perturbed input; the fixture keeps the perturbed input, the
outputs, the loss and the gradients (norms of all, full copies of the small ones, 8x8 slices of the 3x3 ones)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden')


def main():
    sys.path.insert(0, REF)
    from models.DIP import get_net
    from utils.downsampler import Downsampler
    from utils.DIP import get_noise

    torch.set_num_threads(1)
    for (pad, up, H, W, seed) in (('zero', 'nearest', 72, 88, 11), ('zero', 'bilinear', 64, 64, 12),
                                  ('reflection', 'nearest', 64, 96, 13)):
        torch.manual_seed(seed)
        net = get_net(32, 'skip', pad, skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5, upsample_mode=up)
        net_input = get_noise(32, 'noise', (H, W)).detach()
        z0 = net_input + net_input.clone().normal_() * 0.05                     # DIP.py:52
        gen = torch.Generator().manual_seed(100 + seed)
        hr = torch.rand(1, 3, H, W, generator=gen)
        ds = Downsampler(n_planes=3, factor=4, kernel_type='lanczos2', phase=0.5, preserve_size=True)
        with torch.no_grad():
            lr_img = ds(hr)
        out_hr = net(z0)                                                        # DIP.py:60
        out_lr = ds(out_hr)                                                     # DIP.py:62
        loss = torch.nn.MSELoss()(out_lr, lr_img)                               # DIP.py:65
        loss.backward()                                                         # DIP.py:68
        grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
        fixture = {
            'pad': pad, 'upsample_mode': up, 'H': H, 'W': W, 'factor': 4, 'seed': seed,
            'keys': list(net.state_dict().keys()),
            'z0': z0, 'lr_img': lr_img, 'out_hr': out_hr.detach(), 'out_lr': out_lr.detach(), 'losses': [float(loss)],
            'grad_norms': {k: float(v.double().norm()) for k, v in grads.items()},
            'grad_full': {k: v for k, v in grads.items() if v.numel() <= 4096},
            'grad_slices': {k: v[:8, :8].clone() for k, v in grads.items() if v.dim() == 4 and v.shape[-1] == 3},
        }
        name = f'step_{pad}_{up}_{H}x{W}.pt'
        torch.save(fixture, os.path.join(OUT, name))
        print(name, 'loss', float(loss), os.path.getsize(os.path.join(OUT, name)))


if __name__ == '__main__':
    main()
