"""Runs the reference's OWN evaluation loop -- eval_GAN.GAN_ISR_Batch_eval of the unmodified checkout staged in
baseline/_ref (eval_GAN.py:21-67: generator inference, PSNR / SSIM, the uint8 conversion and the PNG writer) -- either
over the reference's own modules (--impl reference: torch eager) or over this repository's drop-in modules (--impl ours:
deep-super-resolution_b200 first on sys.path, so models.GAN.generator, utils.common and the torchmetrics names resolve
to the B200 library while eval_GAN.py and dataset.py stay the reference's files); with --impl ours it then runs
dsr_b200.GAN_ISR_Batch_eval (device-side uint8 conversion) over the same images and compares the PNG files byte for
byte.  LPIPS (pretrained AlexNet, no weights offline) is substituted outside the reference's code.  One JSON line.

    python tools/run_reference_gan_eval.py --impl ours --device cuda --images 2 --lr-size 20
"""
import argparse
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
PKG = os.path.join(ROOT, 'deep-super-resolution_b200')


def png_digests(out_dir, names):
    return [hashlib.sha256(open(os.path.join(out_dir, 'images', f'{n}.png'), 'rb').read()).hexdigest() for n in names]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--device', default='cuda')
    ap.add_argument('--images', type=int, default=2)
    ap.add_argument('--lr-size', type=int, default=20)
    ap.add_argument('--seed', type=int, default=5)
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
    torch.backends.cudnn.allow_tf32 = False
    import eval_GAN as EG                                     # the reference's eval_GAN.py, unmodified
    import utils.common as UC
    from models.GAN.generator import Generator
    assert os.path.realpath(EG.__file__).startswith(os.path.realpath(REF))
    origin = os.path.realpath(sys.modules['models.GAN.generator'].__file__)
    assert origin.startswith(os.path.realpath(PKG if args.impl == 'ours' else REF)), origin
    common_origin = os.path.realpath(UC.__file__)

    class _NoLpips:                                           # LPIPS = pretrained AlexNet, third party, no weights offline
        def __init__(self, *a, **k): pass
        def to(self, d): return self
        def __call__(self, a, b): return torch.zeros(())
    EG.LPIPS = _NoLpips
    from oracle import gan_oracle as GO                       # harness only: synthetic LR / HR pairs
    dev = torch.device(args.device)
    torch.manual_seed(args.seed)
    gan_G = Generator(factor=8).to(dev)
    gan_G.eval()                                              # eval_GAN.py:94
    h = args.lr_size
    loader = []
    for i in range(args.images):
        g = torch.Generator().manual_seed(args.seed * 100 + i)
        hr = torch.rand(1, 3, 8 * h, 8 * (h + 4 * i), generator=g)
        lr = torch.nn.functional.avg_pool2d(hr, 8)
        loader.append((lr, hr, [f'img{i}']))
    names = [f'img{i}' for i in range(args.images)]
    out = {'impl': args.impl, 'device': args.device, 'driver': 'reference eval_GAN.GAN_ISR_Batch_eval (baseline/_ref/eval_GAN.py)',
           'modules': origin.replace(ROOT + '/', ''), 'utils_common': common_origin.replace(ROOT + '/', '')}
    with tempfile.TemporaryDirectory() as d_ref:
        with torch.no_grad():
            m = EG.GAN_ISR_Batch_eval(gan_G, loader, d_ref, args.images, dev)
        out['metrics'] = {k: float(v) for k, v in m.items()}
        out['png_sha256'] = png_digests(d_ref, names)
    if args.impl == 'ours':
        import dsr_b200
        with tempfile.TemporaryDirectory() as d_ours:
            m2 = dsr_b200.GAN_ISR_Batch_eval(gan_G, loader, d_ours, args.images, dev)
            out['mirror_metrics'] = {k: (float(v) if v is not None else None) for k, v in m2.items()}
            out['mirror_png_sha256'] = png_digests(d_ours, names)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
