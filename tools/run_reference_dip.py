"""Runs the reference's OWN per-image driver -- DIP.DIP_ISR of the unmodified checkout staged in baseline/_ref
(DIP.py:22-123) -- either over the reference's own modules (--impl reference: torch eager, CPU or CUDA) or over this
repository's drop-in modules (--impl ours: deep-super-resolution_b200 first on sys.path, so `models.DIP`,
`utils.downsampler`, `utils.DIP` and the torchmetrics names resolve to the B200 library while DIP.py, dataset.py and
utils/common.py stay the reference's files).  Prints one JSON line.

    python tools/run_reference_dip.py --impl ours --device cuda --size 256 --iters 100
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
PKG = os.path.join(ROOT, 'deep-super-resolution_b200')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--device', default='cuda')
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--iters', type=int, default=100)
    ap.add_argument('--log-freq', type=int, default=25)
    ap.add_argument('--image', type=int, default=0)
    args = ap.parse_args()
    if not os.path.isdir(REF):
        print(json.dumps({'unavailable': 'baseline/_ref is missing (run __graft_entry__.build() where /root/reference exists)'}))
        return
    if args.impl == 'ours':
        sys.path[:0] = [os.path.join(PKG, 'metrics_dropin'), PKG, REF, ROOT]
    else:
        sys.path[:0] = [os.path.join(ROOT, 'tests', 'shims'), REF, ROOT]
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False          # reference arm on CUDA: fp32 eager, TF32 off
    import DIP as ref_dip                             # the reference's DIP.py, unmodified
    from models.DIP import get_net                    # ours or the reference's, by sys.path order
    from torchmetrics.image import PeakSignalNoiseRatio as PSNR, StructuralSimilarityIndexMeasure as SSIM
    assert os.path.realpath(ref_dip.__file__).startswith(os.path.realpath(REF))
    origin = os.path.realpath(sys.modules['models.DIP'].__file__)
    assert origin.startswith(os.path.realpath(PKG if args.impl == 'ours' else REF)), origin

    from oracle import dip_oracle as O               # harness only: the synthetic image pair of SURVEY 8d
    lr_img, hr = O.synthetic_pair(args.image, args.size)
    dev = torch.device(args.device)
    if dev.type == 'cpu':
        torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(args.image)
    net = get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                  upsample_mode='bilinear').to(dev)                                   # DIP.py:169-174
    cfg = {'learning_rate': 0.01, 'reg_noise_std': 0.05, 'num_iter': args.iters}      # DIP.py:316-324
    psnr, ssim = PSNR().to(dev), SSIM(data_range=1.).to(dev)                          # DIP.py:157-158
    lpips = lambda a, b: torch.zeros(())                                              # noqa: E731  (needs AlexNet weights)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    t0 = time.time()
    resolved, metrics = ref_dip.DIP_ISR(net, lr_img, hr, 4, cfg, args.log_freq, psnr, ssim, lpips, dev)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    dt = time.time() - t0
    final = O.psnr(resolved.detach().cpu(), hr.unsqueeze(0))
    print(json.dumps({'impl': args.impl, 'device': args.device, 'driver': 'reference DIP.DIP_ISR (baseline/_ref/DIP.py)',
                      'modules': origin.replace(ROOT + '/', ''), 'size': args.size, 'iters': args.iters,
                      'it_per_s': args.iters / dt, 'seconds': dt, 'psnrs': metrics['psnrs'], 'ssims': metrics['ssims'],
                      'final_psnr': final, 'resolved_shape': list(resolved.shape)}))


if __name__ == '__main__':
    main()
