"""Long-run DIP reference curves (build container only: imports the UNMODIFIED reference from /root/reference).

    python oracle/make_golden_psnr.py [--size 128] [--images 8] [--iters 400]

For image i (oracle.synthetic_pair(i, size), factor 4): torch.manual_seed(i); reference get_net / get_noise /
Downsampler / optimize with a closure repeating DIP.py:47-69 (CPU, fp32).  Records the PSNR (10 log10(1/MSE) on
[0,1]) of net(net_input) against the HR image every 50 iterations and the mean of the last 50, plus the
loss curve.  A second run of image 0 with a different noise seed gives the reference's own seed-to-seed spread.
Output: tests/golden/psnr_<size>.pt
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/root/reference'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--images', type=int, default=8)
    ap.add_argument('--iters', type=int, default=400)
    ap.add_argument('--extend', action='store_true', help='add images [len(existing), --images) to the existing fixture')
    args = ap.parse_args()
    from oracle import dip_oracle as O
    sys.path.insert(0, REF)
    from models.DIP import get_net
    from utils.downsampler import Downsampler
    from utils.DIP import get_noise, get_params, optimize

    torch.set_num_threads(os.cpu_count() or 1)
    out = {'size': args.size, 'iters': args.iters, 'factor': 4, 'lr': 0.01, 'reg_noise_std': 0.05, 'runs': []}
    runs = [(i, i) for i in range(args.images)] + [(0, 12345)]        # (image, seed); last = spread probe
    path = os.path.join(ROOT, 'tests', 'golden', f'psnr_{args.size}.pt')
    if args.extend:                                                   # keep what exists, run the missing images only
        out = torch.load(path)
        assert out['iters'] == args.iters
        have = {(r['image'], r['seed']) for r in out['runs']}
        runs = [r for r in runs if r not in have]
    for img, seed in runs:
        lr_img, hr = O.synthetic_pair(img, args.size)
        torch.manual_seed(seed)
        net = get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                      upsample_mode='bilinear')
        ds = Downsampler(n_planes=3, factor=4, kernel_type='lanczos2', phase=0.5, preserve_size=True)
        net_input = get_noise(32, 'noise', (args.size, args.size)).detach()
        saved, noise = net_input.detach().clone(), net_input.detach().clone()
        target, hr4 = lr_img.unsqueeze(0), hr.unsqueeze(0)
        mse = torch.nn.MSELoss()
        rec = {'image': img, 'seed': seed, 'psnr': [], 'loss': [], 'it': 0}
        t0 = time.time()

        def closure():
            z = saved + noise.normal_() * 0.05                      # DIP.py:52
            o = net(z)                                              # DIP.py:60
            loss = mse(ds(o), target)                               # DIP.py:62-65
            loss.backward()                                         # DIP.py:68
            rec['loss'].append(float(loss))
            rec['psnr'].append(O.psnr(o.detach(), hr4))
            rec['it'] += 1
            return loss

        optimize('adam', get_params('net', net, net_input), closure, 0.01, args.iters)
        p = torch.tensor(rec['psnr'])
        rec['psnr_last50'] = float(p[-50:].mean())
        rec['psnr_every50'] = [float(v) for v in p[49::50]]
        rec['loss_every50'] = [rec['loss'][k] for k in range(49, args.iters, 50)]
        del rec['psnr'], rec['loss']
        out['runs'].append(rec)
        print(f"image {img} seed {seed}: PSNR(last 50) {rec['psnr_last50']:.3f} dB  loss {rec['loss_every50'][-1]:.3e} "
              f"({time.time() - t0:.0f} s)", flush=True)
    main_runs = [r for r in out['runs'] if r['image'] == r['seed']]
    probe = [r for r in out['runs'] if r['image'] != r['seed']][0]
    first = [r for r in main_runs if r['image'] == probe['image']][0]
    out['mean_psnr_last50'] = sum(r['psnr_last50'] for r in main_runs) / len(main_runs)
    out['seed_spread_image0'] = abs(probe['psnr_last50'] - first['psnr_last50'])
    print('mean', out['mean_psnr_last50'], 'spread(image 0, other seed)', out['seed_spread_image0'])
    torch.save(out, path)


if __name__ == '__main__':
    main()
