"""Runs the reference's OWN training loop -- train_GAN.GAN_ISR_train of the unmodified checkout staged in baseline/_ref
(train_GAN.py:22-136, i.e. do_epoch with its loss.backward() / torch.optim.Adam) -- either over the reference's own
modules (--impl reference: torch eager, CPU or CUDA, fp32, TF32 off) or over this repository's drop-in modules (--impl
ours: deep-super-resolution_b200 first on sys.path, so models.GAN.generator / models.GAN.discriminator / utils.GAN and
the torchmetrics names resolve to the B200 library while train_GAN.py, dataset.py and utils/common.py stay the
reference's files).  The VGG19 of the perceptual loss has random weights in both arms (same seed; no pretrained file
exists offline).  Prints one JSON line.

    python tools/run_reference_gan.py --impl ours --device cuda --batch 8 --lr-size 24 --epochs 3
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
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--lr-size', type=int, default=24)
    ap.add_argument('--epochs', type=int, default=1)
    ap.add_argument('--lr', type=float, default=1e-4)
    ap.add_argument('--seed', type=int, default=31)
    args = ap.parse_args()
    if not os.path.isdir(REF):
        print(json.dumps({'unavailable': 'baseline/_ref is missing (run __graft_entry__.build() where /root/reference exists)'}))
        return
    os.environ['DSR_VGG_RANDOM'] = '1'                 # no pretrained VGG19 offline: do not even try to download
    if args.impl == 'ours':
        sys.path[:0] = [os.path.join(PKG, 'metrics_dropin'), PKG, REF, ROOT]
    else:
        sys.path[:0] = [os.path.join(ROOT, 'tests', 'shims'), REF, ROOT]
    import torch
    import torchvision
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    import train_GAN as TG                                    # the reference's train_GAN.py, unmodified
    import utils.GAN as UG
    from models.GAN.generator import Generator
    from models.GAN.discriminator import Discriminator
    assert os.path.realpath(TG.__file__).startswith(os.path.realpath(REF))
    origin = os.path.realpath(sys.modules['models.GAN.discriminator'].__file__)
    assert origin.startswith(os.path.realpath(PKG if args.impl == 'ours' else REF)), origin
    class _NoLpips:                                          # LPIPS = pretrained AlexNet, third party, no weights offline:
        def __init__(self, *a, **k): pass                    # substituted OUTSIDE the reference's code in both arms
        def to(self, d): return self
        def __call__(self, a, b): return torch.zeros(())
    TG.LPIPS = _NoLpips
    if args.impl == 'reference':
        UG.vgg19 = lambda weights=None, **kw: torchvision.models.vgg19(weights=None)     # noqa: E731
    from oracle import gan_train_oracle as O                 # harness only: the synthetic LR / HR batch
    dev = torch.device(args.device)
    if dev.type == 'cpu':
        torch.set_num_threads(os.cpu_count() or 1)
    f = 8
    torch.manual_seed(args.seed)
    gan_G = Generator(factor=f).to(dev)                      # train_GAN.py:155-164
    gan_D = Discriminator((args.lr_size * f, args.lr_size * f)).to(dev)
    gan_G.train(); gan_D.train()
    LR, HR = O.synthetic_batch(args.seed + 7, args.batch, (args.lr_size, args.lr_size), f)
    loader = [(LR, HR, 0)]
    torch.manual_seed(args.seed + 1000)                      # the VGG19 initialisation drawn inside GAN_ISR_train
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    t0 = time.time()
    G, D, metrics = TG.GAN_ISR_train(gan_G, gan_D, args.lr, loader, args.epochs, 10 ** 9, dev)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    dt = time.time() - t0
    sdG = G.state_dict()
    # train_GAN.py:128-129 stores the final losses under swapped labels
    print(json.dumps({'impl': args.impl, 'device': args.device, 'driver': 'reference train_GAN.GAN_ISR_train (baseline/_ref)',
                      'modules': origin.replace(ROOT + '/', ''), 'batch': args.batch, 'lr_size': args.lr_size,
                      'epochs': args.epochs, 'seconds': dt, 'steps_per_s': args.epochs / dt,
                      'loss_D': metrics['Final Generator loss'], 'loss_G': metrics['Final Discriminator loss'],
                      'psnr0': metrics['Average PSNR during training'][0],
                      'bn_mean_abs': float(sdG['bn1.running_mean'].abs().mean()),
                      'conv3_w_abs': float(sdG['conv3.weight'].abs().mean())}))


if __name__ == '__main__':
    main()
