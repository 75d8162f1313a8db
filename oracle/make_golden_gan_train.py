"""Generate tests/golden/gan_train_*.pt by EXECUTING THE UNMODIFIED REFERENCE training step (build container only).

    python oracle/make_golden_gan_train.py

``train_GAN.GAN_ISR_train`` (train_GAN.py:22-136) is imported from /root/reference and run for one epoch over a
one-batch loader, i.e. exactly one ``do_epoch`` (train_GAN.py:38-71) with its two Adam steps.  Two things the offline
container cannot provide are substituted OUTSIDE the reference's code: ``torchvision.models.vgg19`` is patched to return
the randomly initialised network under a recorded seed (``Vgg19Loss.__init__`` would download IMAGENET1K_V1), and
``torchmetrics`` comes from tests/shims.  Recorded per case: the seed recipe, the LR / HR batch, both losses, per-key
checksums of the initial and post-step state dicts of G and D, the generator's ``.grad`` (what ``loss_G.backward()`` left)
and a strided sample of every post-step tensor.
"""
import os
import sys

import torch

REF = '/root/reference'
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
OUT = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'shims'))

CASES = [  # name, batch, LR h, w, factor, seed, lr
    ('gan_train_2x8x8', 2, 8, 8, 8, 21, 1e-4),
    ('gan_train_3x16x8', 3, 16, 8, 8, 22, 1e-3),
]


def checksum(t):
    t = t.detach().double().flatten()
    return (float(t.sum()), float(t.abs().sum()))


def sample(t, n=64):
    f = t.detach().flatten()
    step = max(1, f.numel() // n)
    return f[::step][:n].clone()


def main():
    sys.path.insert(0, REF)
    import torchvision
    import utils.GAN as UG                                   # noqa: E402  (the unmodified reference)
    import train_GAN as TG                                   # noqa: E402
    from models.GAN.generator import Generator              # noqa: E402
    from models.GAN.discriminator import Discriminator      # noqa: E402
    from oracle import gan_train_oracle as O
    torch.set_num_threads(1)
    for name, b, h, w, factor, seed, lr in CASES:
        torch.manual_seed(seed)
        gan_G = Generator(factor=factor)
        gan_D = Discriminator((h * factor, w * factor))
        gan_G.train(); gan_D.train()
        init = {'G': {k: checksum(v) for k, v in gan_G.state_dict().items()},
                'D': {k: checksum(v) for k, v in gan_D.state_dict().items()}}

        def random_vgg19(weights=None, **kw):
            torch.manual_seed(seed + 1000)
            return torchvision.models.vgg19(weights=None)
        UG.vgg19 = random_vgg19
        LR, HR = O.synthetic_batch(seed + 7, b, (h, w), factor)
        loader = [(LR, HR, 0)]
        losses = {}
        orig_get_loss_D = UG.get_loss_D

        # the losses are only returned as python floats through prints; capture them from the metric dict below
        G, D, metrics = TG.GAN_ISR_train(gan_G, gan_D, lr, loader, 1, 1, torch.device('cpu'))
        # train_GAN.py:128-129 stores them under swapped labels: 'Final Generator loss' holds loss_D and vice versa
        losses['loss_D'] = metrics['Final Generator loss']
        losses['loss_G'] = metrics['Final Discriminator loss']
        assert UG.get_loss_D is orig_get_loss_D
        fx = dict(batch=b, lr_hw=(h, w), factor=factor, seed=seed, vgg_seed=seed + 1000, lr=lr, LR=LR, HR=HR, **losses,
                  init=init,
                  post={'G': {k: checksum(v) for k, v in G.state_dict().items()},
                        'D': {k: checksum(v) for k, v in D.state_dict().items()}},
                  post_sample={'G': {k: sample(v) for k, v in G.state_dict().items()},
                               'D': {k: sample(v) for k, v in D.state_dict().items()}},
                  grad_G={k: (checksum(p.grad), sample(p.grad)) for k, p in G.named_parameters() if p.grad is not None})
        torch.save(fx, os.path.join(OUT, name + '.pt'))
        print(name, losses, len(fx['grad_G']))


if __name__ == '__main__':
    main()
