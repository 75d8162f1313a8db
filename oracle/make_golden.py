"""Generate tests/golden/*.pt by EXECUTING THE UNMODIFIED REFERENCE (build container only).

Run here, where /root/reference exists:   python oracle/make_golden.py
The GPU box has no /root/reference; tests only read the committed fixtures.

What is recorded (all produced by the reference's own modules -- models/DIP, utils/downsampler.py,
utils/DIP.py -- driven by a closure that repeats DIP.py:47-69 line for line, minus the
torchmetrics logging branch, which is out of scope and not installed):

* lanczos.pt      get_kernel tables for factor 4/8/16 (float64) and Downsampler outputs on a
                  fixed input.
* init_seedS.pt   per-key float64 checksums of get_net(...)'s same-seed parameters (the tensors
                  themselves are 8.9 MB; the oracle regenerates them and must hit the checksums).
* step_HxW.pt     one teacher-forced step: z, out_HR, out_LR, loss, per-key gradient norms and
                  leading values, selected full gradients, post-Adam parameter checksums,
                  BN running statistics, and the loss trajectory of 3 iterations of
                  utils.DIP.optimize.
"""
import os
import sys

import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def checksum(t):
    t = t.detach().double().flatten()
    return (float(t.sum()), float(t.abs().sum()), float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64)).sum() / max(1, t.numel())))


def main():
    sys.path.insert(0, REF)
    from models.DIP import get_net                      # noqa: E402
    from utils.downsampler import Downsampler, get_kernel  # noqa: E402
    from utils.DIP import get_noise, get_params, optimize  # noqa: E402

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)   # fixes the reduction order of the fixtures

    # ---- Lanczos tables and downsampler outputs -------------------------------------------
    g = torch.Generator().manual_seed(7)
    lan = {}
    for f in (4, 8, 16):
        lan[f'kernel_f{f}'] = torch.from_numpy(get_kernel(f, 'lanczos', 0.5, 4 * f + 1, support=2))
    x = torch.rand(1, 3, 64, 96, generator=g)
    for f in (4, 8, 16):
        ds = Downsampler(n_planes=3, factor=f, kernel_type='lanczos2', phase=0.5, preserve_size=True)
        xin = x.clone().requires_grad_(True)
        y = ds(xin)
        gy = torch.rand(y.shape, generator=g)
        (y * gy).sum().backward()
        lan[f'y_f{f}'] = y.detach()
        lan[f'gy_f{f}'] = gy
        lan[f'gx_f{f}'] = xin.grad.detach()
    lan['x'] = x
    torch.save(lan, os.path.join(OUT, 'lanczos.pt'))

    # ---- same-seed initialisation -----------------------------------------------------------
    def make_net():
        return get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4,
                       num_scales=5, upsample_mode='bilinear')

    for seed in (0, 3):
        torch.manual_seed(seed)
        net = make_net()
        sd = net.state_dict()
        torch.save({'keys': list(sd.keys()),
                    'shapes': {k: tuple(v.shape) for k, v in sd.items()},
                    'checksums': {k: checksum(v) for k, v in sd.items() if v.dtype.is_floating_point},
                    'param_order': [n for n, _ in net.named_parameters()]},
                   os.path.join(OUT, f'init_seed{seed}.pt'))

    # ---- one teacher-forced step + a 3-iteration trajectory ---------------------------------
    for (H, W, factor, seed) in ((64, 64, 4, 0), (64, 96, 4, 3), (72, 88, 4, 5)):   # last: odd level sizes + Concat crop
        torch.manual_seed(seed)
        net = make_net()
        net_input = get_noise(32, 'noise', (H, W)).detach()
        net_input_saved = net_input.detach().clone()
        noise = net_input.detach().clone()
        gen = torch.Generator().manual_seed(100 + seed)
        hr = torch.rand(1, 3, H, W, generator=gen)
        ds = Downsampler(n_planes=3, factor=factor, kernel_type='lanczos2', phase=0.5, preserve_size=True)
        with torch.no_grad():
            lr_img = ds(hr)
        msef = torch.nn.MSELoss()
        reg = 0.05
        rec = {'losses': [], 'z': []}

        def closure():
            z = net_input_saved + (noise.normal_() * reg)          # DIP.py:52
            rec['z'].append(z.clone())
            out_hr = net(z)                                         # DIP.py:60
            out_lr = ds(out_hr)                                     # DIP.py:62
            loss = msef(out_lr, lr_img)                             # DIP.py:65
            loss.backward()                                         # DIP.py:68
            if len(rec['losses']) == 0:
                rec['out_hr'] = out_hr.detach().clone()
                rec['out_lr'] = out_lr.detach().clone()
                rec['grads'] = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
            rec['losses'].append(float(loss))
            return loss

        params = get_params('net', net, net_input)
        # iteration 1 by hand so that post-step state can be captured, then 2 more via optimize
        optimize('adam', params, closure, 0.01, 1)
        sd1 = {k: v.detach().clone() for k, v in net.state_dict().items()}
        optimize('adam', params, closure, 0.01, 2)   # NB: fresh Adam state, as a second DIP_ISR call would

        grads = rec['grads']
        keep_full = [k for k in grads if grads[k].numel() <= 4096 or k in ('9.1.weight', '1.0.1.1.weight')]
        fixture = {
            'H': H, 'W': W, 'factor': factor, 'seed': seed, 'reg_noise_std': reg, 'lr': 0.01,
            'net_input_saved_checksum': checksum(net_input_saved),
            'z0': rec['z'][0], 'hr': hr, 'lr_img': lr_img,
            'out_hr': rec['out_hr'], 'out_lr': rec['out_lr'], 'losses': rec['losses'],
            'grad_norms': {k: float(v.double().norm()) for k, v in grads.items()},
            'grad_head': {k: v.flatten()[:8].clone() for k, v in grads.items()},
            'grad_full': {k: grads[k] for k in keep_full},
            # a 3x3 conv gradient slice, to pin wgrad layouts: [co 0..7, ci 0..7, :, :]
            'grad_slices': {k: grads[k][:8, :8].clone() for k in grads if grads[k].dim() == 4 and grads[k].shape[-1] == 3},
            'post_adam_checksums': {k: checksum(v) for k, v in sd1.items() if v.dtype.is_floating_point},
            'post_adam_small': {k: v for k, v in sd1.items() if v.numel() <= 132},
        }
        torch.save(fixture, os.path.join(OUT, f'step_{H}x{W}.pt'))
        print(f'step_{H}x{W}: losses {rec["losses"]}')

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
