"""CPU oracle for the SRResNet generator (inference).  TEST INFRASTRUCTURE ONLY.

Checker for ``dsr_b200.gan.Generator`` (BASELINE configs[3]).  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU arm of ``tools/gan_bench.py`` may import it.

Functional restatement (plain fp32 ``torch.nn.functional`` calls on a state dict, no ``nn.Module`` graph) of
``models/GAN/generator.py`` of the reference in EVAL mode, as ``eval_GAN.py:87-94`` runs it: every function cites the
reference lines it follows.  Pinned against outputs of the reference itself: ``oracle/make_golden_gan.py`` imports the
unmodified ``models.GAN.generator.Generator`` from ``/root/reference`` (build container only) and writes
``tests/golden/gan_*.pt``; ``tests/test_gan_oracle.py`` reproduces them.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5                      # nn.BatchNorm2d default (generator.py:8,12,54)
N_FEAT = 64                        # generator.py:7,47
SHUFFLES = {8: 3, 16: 4}           # generator.py:55-58 (no other factor is constructible)


def init_state_dict(factor: int = 8, residual_blocks: int = 16) -> Dict[str, Tensor]:
    """Fresh parameters + buffers with the reference's initialisation and RNG consumption order.

    ``Generator.__init__`` (generator.py:45-66) constructs conv1 (9x9, 3->64), the residual blocks (each conv1,
    conv2: generator.py:7-12), conv2 (3x3), the shuffle blocks (conv 64->256, generator.py:30) and conv3 (9x9,
    64->3), in that order; only ``nn.Conv2d`` draws random numbers (weight then bias), ``nn.PReLU`` starts at 0.25
    and ``nn.BatchNorm2d`` at (1, 0, 0, 1).  Call under ``torch.manual_seed(s)`` to get the reference's same-seed
    weights bit for bit."""
    sd: Dict[str, Tensor] = {}

    def conv(name, cin, cout, k):
        m = torch.nn.Conv2d(cin, cout, k)
        sd[name + '.weight'] = m.weight.detach().clone()
        sd[name + '.bias'] = m.bias.detach().clone()

    def bn(name):
        sd[name + '.weight'] = torch.ones(N_FEAT)
        sd[name + '.bias'] = torch.zeros(N_FEAT)
        sd[name + '.running_mean'] = torch.zeros(N_FEAT)
        sd[name + '.running_var'] = torch.ones(N_FEAT)
        sd[name + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    def prelu(name):
        sd[name + '.weight'] = torch.full((1,), 0.25)

    conv('conv1', 3, N_FEAT, 9)
    prelu('prelu1')
    for i in range(residual_blocks):
        p = f'residual_blocks.{i}.'
        conv(p + 'conv1', N_FEAT, N_FEAT, 3)
        bn(p + 'bn1')
        prelu(p + 'prelu1')
        conv(p + 'conv2', N_FEAT, N_FEAT, 3)
        bn(p + 'bn2')
    conv('conv2', N_FEAT, N_FEAT, 3)
    bn('bn1')
    for i in range(SHUFFLES[factor]):
        p = f'pixel_shuffle_blocks.{i}.'
        conv(p + 'conv1', N_FEAT, 4 * N_FEAT, 3)
        prelu(p + 'prelu1')
    conv('conv3', N_FEAT, 3, 9)
    return sd


def perturb_trained_state(sd: Dict[str, Tensor], seed: int) -> None:
    """Makes a fresh state dict look TRAINED (in place, deterministic): non-trivial BatchNorm affine parameters and
    running statistics and distinct PReLU slopes, so that eval-mode BatchNorm and PReLU are actually exercised (a
    fresh net has identity BatchNorm in eval mode).  Applied identically to the reference module by
    make_golden_gan.py (through load_state_dict) and by the tests."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        t = sd[k]
        if k.endswith('running_mean'):
            t.copy_(torch.randn(t.shape, generator=g) * 0.2)
        elif k.endswith('running_var'):
            t.copy_(torch.rand(t.shape, generator=g) * 1.5 + 0.25)
        elif 'bn' in k and k.endswith('.weight'):
            t.copy_(torch.rand(t.shape, generator=g) * 0.6 + 0.5)
        elif 'bn' in k and k.endswith('.bias'):
            t.copy_(torch.randn(t.shape, generator=g) * 0.1)
        elif 'prelu' in k:
            t.copy_(torch.rand(t.shape, generator=g) * 0.4 + 0.05)


def _bn_eval(x: Tensor, sd: Dict[str, Tensor], name: str) -> Tensor:
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'], sd[name + '.weight'],
                        sd[name + '.bias'], training=False, momentum=0.1, eps=BN_EPS)


def generator_forward(sd: Dict[str, Tensor], x: Tensor, factor: int = 8, residual_blocks: int = 16,
                      record: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """``Generator.forward`` (generator.py:68-81) with BatchNorm in eval mode (eval_GAN.py:94).
    x: [B, 3, h, w] in [0, 1] -> [B, 3, factor*h, factor*w] in (-1, 1).  ``record`` (tests) receives the
    intermediates x0, block{i}_t, block{i}, trunk, s{i}."""
    rec = record if record is not None else {}
    z = F.conv2d(x, sd['conv1.weight'], sd['conv1.bias'], padding=4)                       # :70
    x0 = F.prelu(z, sd['prelu1.weight'])                                                   # :71
    rec['x0'] = x0
    z = x0
    for i in range(residual_blocks):                                                      # :72, ResidualBlock.forward :14-25
        p = f'residual_blocks.{i}.'
        t = F.conv2d(z, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'], padding=1)
        t = F.prelu(_bn_eval(t, sd, p + 'bn1'), sd[p + 'prelu1.weight'])
        rec[f'block{i}_t'] = t
        t = F.conv2d(t, sd[p + 'conv2.weight'], sd[p + 'conv2.bias'], padding=1)
        z = z + _bn_eval(t, sd, p + 'bn2')
        rec[f'block{i}'] = z
    z = F.conv2d(z, sd['conv2.weight'], sd['conv2.bias'], padding=1)                       # :73
    z = x0 + _bn_eval(z, sd, 'bn1')                                                        # :74-76
    rec['trunk'] = z
    for i in range(SHUFFLES[factor]):                                                      # :78, PixelShuffleBlock.forward :36-41
        p = f'pixel_shuffle_blocks.{i}.'
        z = F.conv2d(z, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'], padding=1)
        z = F.prelu(F.pixel_shuffle(z, 2), sd[p + 'prelu1.weight'])
        rec[f's{i}'] = z
    z = F.conv2d(z, sd['conv3.weight'], sd['conv3.bias'], padding=4)                       # :80
    return torch.tanh(z)                                                                   # :82


def flops_per_image(h: int, w: int, factor: int = 8, residual_blocks: int = 16) -> float:
    """Multiply-add FLOPs (2 per MAC) of one forward on an h x w input (98.13 GFLOP at 96 x 96, x8)."""
    px = h * w
    f = 2.0 * px * N_FEAT * 3 * 81
    f += (2 * residual_blocks + 1) * 2.0 * px * N_FEAT * N_FEAT * 9
    for _ in range(SHUFFLES[factor]):
        f += 2.0 * px * 4 * N_FEAT * N_FEAT * 9
        px *= 4
    f += 2.0 * px * 3 * N_FEAT * 81
    return f
