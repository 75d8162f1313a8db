"""Generate tests/golden/gan_*.pt by EXECUTING THE UNMODIFIED REFERENCE generator (build container only).

    python oracle/make_golden_gan.py

``models.GAN.generator.Generator`` is imported from /root/reference, constructed under a fixed seed, given
"trained-looking" BatchNorm statistics / PReLU slopes (oracle.gan_oracle.perturb_trained_state, through
load_state_dict) and evaluated in eval mode as eval_GAN.py:87-94 does.  Recorded per case: the seed recipe, per-key
checksums of the state dict (the tensors are ~7 MB; the oracle regenerates them and must hit the checksums), the input
batch and the reference's output.
"""
import os
import sys

import torch

REF = '/root/reference'
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
OUT = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)

CASES = [  # name, factor, batch, h, w, init seed, perturbation seed
    ('gan_f8_2x20x24', 8, 2, 20, 24, 11, 5),
    ('gan_f8_1x17x23', 8, 1, 17, 23, 12, 6),       # sizes that are no multiple of the 8 x 16 pixel tiles
    ('gan_f16_1x16x16', 16, 1, 16, 16, 13, 7),
]


def checksum(t):
    t = t.detach().double().flatten()
    return (float(t.sum()), float(t.abs().sum()))


def main():
    sys.path.insert(0, REF)
    from models.GAN.generator import Generator      # noqa: E402  (the unmodified reference)
    from oracle import gan_oracle as go
    torch.set_num_threads(1)
    for name, factor, b, h, w, seed, pseed in CASES:
        torch.manual_seed(seed)
        net = Generator(factor=factor)
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        go.perturb_trained_state(sd, pseed)
        net.load_state_dict(sd)
        net.eval()
        g = torch.Generator().manual_seed(seed + 100)
        x = torch.rand(b, 3, h, w, generator=g)
        with torch.no_grad():
            y = net(x)
        fx = dict(factor=factor, seed=seed, perturb_seed=pseed, x=x, y=y,
                  keys=list(sd.keys()), checksums={k: checksum(v) for k, v in sd.items()})
        torch.save(fx, os.path.join(OUT, name + '.pt'))
        print(name, tuple(y.shape), float(y.abs().mean()))


if __name__ == '__main__':
    main()
