"""Oracle for the SRGAN TRAINING step.  TEST INFRASTRUCTURE ONLY.

Checker for ``dsr_b200.gan_train`` (BASELINE configs[4]).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import it; the product never does.

Functional restatement (plain fp32 ``torch.nn.functional`` calls on state dicts, gradients through torch autograd; runs
on CPU, or on a CUDA device with TF32 off when a test wants the full-size case in seconds) of:

* ``Generator`` in TRAIN mode          -- models/GAN/generator.py:14-25, 36-41, 68-81
* ``Discriminator``                     -- models/GAN/discriminator.py:14-19, 57-74
* ``Vgg19Loss`` incl. the transform     -- utils/GAN.py:62-88 (torchvision ``ImageClassification``: resize 256 bilinear,
                                           centre crop 224, normalise; VGG19 ``features[:36]``)
* ``get_loss_D`` / ``PerceptualLoss``   -- utils/GAN.py:96-123
* ``do_epoch`` with both Adam steps     -- train_GAN.py:38-71, torch.optim.Adam defaults (train_GAN.py:34-35)

Pinned against the reference itself: ``oracle/make_golden_gan_train.py`` executes the UNMODIFIED ``train_GAN.
GAN_ISR_train`` (one batch, one epoch; torchvision's ``vgg19`` patched to random weights because no pretrained file
exists offline, torchmetrics through the test shim) and stores losses, post-step parameter checksums / slices and the
generator gradients in ``tests/golden/gan_train_*.pt``; ``tests/test_gan_train_oracle.py`` reproduces them.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SHUFFLES = {8: 3, 16: 4}
VGG_CFG = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512)
VGG_MEAN = (0.485, 0.456, 0.406)
VGG_STD = (0.229, 0.224, 0.225)
D_BLOCKS = ((64, 64, 2), (64, 128, 1), (128, 128, 2), (128, 256, 1), (256, 256, 2), (256, 512, 1), (512, 512, 2))


# ---------------------------------------------------------------------------------------------
# same-seed initialisation (construction order of the reference modules)
# ---------------------------------------------------------------------------------------------
def _conv(sd, name, cin, cout, k):
    m = torch.nn.Conv2d(cin, cout, k)
    sd[name + '.weight'] = m.weight.detach().clone()
    sd[name + '.bias'] = m.bias.detach().clone()


def _bn(sd, name, c):
    sd[name + '.weight'] = torch.ones(c)
    sd[name + '.bias'] = torch.zeros(c)
    sd[name + '.running_mean'] = torch.zeros(c)
    sd[name + '.running_var'] = torch.ones(c)
    sd[name + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)


def init_discriminator(hr_shape: Tuple[int, int]) -> Dict[str, Tensor]:
    """discriminator.py:22-45: conv, the seven blocks (conv1, bn1), dense1, dense2 -- only Conv2d / Linear draw."""
    sd: Dict[str, Tensor] = {}
    _conv(sd, 'conv', 3, 64, 3)
    for i, (a, b, _) in enumerate(D_BLOCKS):
        _conv(sd, f'convblocks.{i}.conv1', a, b, 3)
        _bn(sd, f'convblocks.{i}.bn1', b)
    # fc_input_shape (discriminator.py:47-55) pushes a batch of ones through conv / convblocks IN TRAIN MODE while the
    # module is being constructed: every BatchNorm's running statistics have already taken one update
    # (num_batches_tracked = 1) before training starts
    with torch.no_grad():
        ns: Dict[str, Tensor] = {}
        hcur = F.leaky_relu(F.conv2d(torch.ones(1, 3, hr_shape[0], hr_shape[1]), sd['conv.weight'], sd['conv.bias'],
                                     padding=1), 0.2)
        for i, (_, _, s) in enumerate(D_BLOCKS):
            p = f'convblocks.{i}.'
            hcur = F.conv2d(hcur, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'], stride=s, padding=1)
            hcur = F.leaky_relu(_bn_train(sd, p + 'bn1', hcur, ns), 0.2)
            sd[p + 'bn1.num_batches_tracked'] = torch.ones((), dtype=torch.long)
        sd.update(ns)
    fin = hcur.reshape(1, -1).shape[1]
    for name, a, b in (('dense1', fin, 1024), ('dense2', 1024, 1)):
        m = torch.nn.Linear(a, b)
        sd[name + '.weight'] = m.weight.detach().clone()
        sd[name + '.bias'] = m.bias.detach().clone()
    return sd


def init_vgg() -> Dict[str, Tensor]:
    """torchvision ``vgg19(weights=None).features[:36]`` wrapped as utils/GAN.py:69 does (keys ``net.0.<idx>.*``)."""
    from torchvision.models import vgg19
    feats = vgg19(weights=None).features[:36]
    return {'net.0.' + k: v.detach().clone() for k, v in feats.state_dict().items()}


def param_keys(sd: Dict[str, Tensor]) -> List[str]:
    return [k for k in sd if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))]


# ---------------------------------------------------------------------------------------------
# precision-class yardstick: the same arithmetic with bf16 rounding at the CUDA path's 16-bit storage points
# ---------------------------------------------------------------------------------------------
def bf16_points():
    """``q`` for generator_train / discriminator_train: rounds to bfloat16 where libdsr_b200 stores 16-bit tensors -- the
    packed conv weights, every convolution input / raw output, every BatchNorm + activation output -- with a
    straight-through gradient (the CUDA backward keeps its gradients in bf16 as well, which this does not model).
    The CUDA path must agree with THIS oracle more closely than with the fp32 one; what is left between the two is the
    accumulation order."""
    def q(t: Tensor) -> Tensor:
        return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()
    return q


def _q(q, t):
    return t if q is None else q(t)


# ---------------------------------------------------------------------------------------------
# forward passes
# ---------------------------------------------------------------------------------------------
def _bn_train(sd, name, x, new_stats):
    """nn.BatchNorm2d in train mode: batch statistics, running statistics updated with momentum 0.1."""
    rm = sd[name + '.running_mean'].clone()
    rv = sd[name + '.running_var'].clone()
    y = F.batch_norm(x, rm, rv, sd[name + '.weight'], sd[name + '.bias'], True, 0.1, 1e-5)
    if new_stats is not None:
        new_stats[name + '.running_mean'] = rm
        new_stats[name + '.running_var'] = rv
    return y


def generator_train(sd: Dict[str, Tensor], x: Tensor, factor: int = 8, blocks: int = 16, new_stats=None, taps=None,
                    q=None) -> Tensor:
    """generator.py:68-81 with ResidualBlock.forward (:14-25) and PixelShuffleBlock.forward (:36-41).  q: optional
    rounding at the 16-bit storage points (bf16_points)."""
    def conv(name, h, pad):
        return _q(q, F.conv2d(_q(q, h), _q(q, sd[name + '.weight']), sd[name + '.bias'], padding=pad))
    z = conv('conv1', x, 4)
    x0 = _q(q, F.prelu(z, sd['prelu1.weight']))
    h = x0
    if taps is not None:
        taps['g_x0'] = x0
    for i in range(blocks):
        p = f'residual_blocks.{i}.'
        t = conv(p + 'conv1', h, 1)
        t = _q(q, F.prelu(_bn_train(sd, p + 'bn1', t, new_stats), sd[p + 'prelu1.weight']))
        t = conv(p + 'conv2', t, 1)
        h = _q(q, h + _bn_train(sd, p + 'bn2', t, new_stats))
        if taps is not None:
            taps[f'g_x{i + 1}'] = h
    t = conv('conv2', h, 1)
    h = _q(q, x0 + _bn_train(sd, 'bn1', t, new_stats))
    if taps is not None:
        taps['g_t'] = h
    for i in range(SHUFFLES[factor]):
        p = f'pixel_shuffle_blocks.{i}.'
        h = conv(p + 'conv1', h, 1)
        h = _q(q, F.prelu(F.pixel_shuffle(h, 2), sd[p + 'prelu1.weight']))
        if taps is not None:
            taps[f'g_u{i}'] = h
    return torch.tanh(F.conv2d(h, _q(q, sd['conv3.weight']), sd['conv3.bias'], padding=4))


def discriminator_train(sd: Dict[str, Tensor], x: Tensor, new_stats=None, taps=None, q=None) -> Tensor:
    """discriminator.py:57-74 with DiscriminatorConvBlock.forward (:14-19).  q: see generator_train."""
    h = _q(q, F.leaky_relu(F.conv2d(_q(q, x), _q(q, sd['conv.weight']), sd['conv.bias'], padding=1), 0.2))
    if taps is not None:
        taps['d_h0'] = h
    for i, (_, _, s) in enumerate(D_BLOCKS):
        p = f'convblocks.{i}.'
        h = _q(q, F.conv2d(h, _q(q, sd[p + 'conv1.weight']), sd[p + 'conv1.bias'], stride=s, padding=1))
        h = _q(q, F.leaky_relu(_bn_train(sd, p + 'bn1', h, new_stats), 0.2))
        if taps is not None:
            taps[f'd_h{i + 1}'] = h
    h = h.reshape(h.shape[0], -1)
    h = F.leaky_relu(F.linear(h, sd['dense1.weight'], sd['dense1.bias']), 0.2)
    return torch.sigmoid(F.linear(h, sd['dense2.weight'], sd['dense2.bias']))


def _bilinear_matrix(n_in: int, n_out: int, dtype, device) -> Tensor:
    """[n_out, n_in] interpolation matrix of ATen's bilinear resize, align_corners=False (with antialias=True the
    triangle filter has support 1 when ENLARGING, which is the same two taps and weights)."""
    scale = n_in / n_out
    src = ((torch.arange(n_out, dtype=torch.float64) + 0.5) * scale - 0.5).clamp_(min=0.0)
    i0 = src.floor().long().clamp_(max=n_in - 1)
    i1 = (i0 + 1).clamp_(max=n_in - 1)
    l1 = (src - i0.double())
    m = torch.zeros(n_out, n_in, dtype=torch.float64)
    m[torch.arange(n_out), i0] += 1.0 - l1
    m[torch.arange(n_out), i1] += l1
    return m.to(dtype=dtype, device=device)


def _vgg_sizes(h: int, w: int):
    if h <= w:
        hr, wr = 256, int(256 * w / h)
    else:
        hr, wr = int(256 * h / w), 256
    return hr, wr, int(round((hr - 224) / 2.0)), int(round((wr - 224) / 2.0))


def _vgg_crop_norm(y: Tensor, top: int, left: int) -> Tensor:
    y = y[:, :, top:top + 224, left:left + 224]
    mean = torch.tensor(VGG_MEAN, dtype=y.dtype, device=y.device).view(1, 3, 1, 1)
    std = torch.tensor(VGG_STD, dtype=y.dtype, device=y.device).view(1, 3, 1, 1)
    return (y - mean) / std


def vgg_transform(x: Tensor) -> Tensor:
    """VGG19_Weights.IMAGENET1K_V1.transforms() (utils/GAN.py:76-77) on a float batch: resize the smaller edge to 256
    (the ATen call torchvision's resize makes: bilinear, align_corners=False, antialias=True), centre crop 224,
    normalise.  Only the enlarging case (patch <= 256) is needed."""
    _, _, h, w = x.shape
    assert min(h, w) <= 256
    hr, wr, top, left = _vgg_sizes(h, w)
    y = F.interpolate(x, size=(hr, wr), mode='bilinear', align_corners=False, antialias=True)
    return _vgg_crop_norm(y, top, left)


def vgg_transform_explicit(x: Tensor) -> Tensor:
    """The same transform with the resize written out as two interpolation matrices -- the arithmetic the CUDA kernels
    g_vgg_pre_fwd_kernel / g_vgg_pre_bwd_kernel implement (checked against torchvision's preset in the tests)."""
    _, _, h, w = x.shape
    hr, wr, top, left = _vgg_sizes(h, w)
    my = _bilinear_matrix(h, hr, x.dtype, x.device)
    mx = _bilinear_matrix(w, wr, x.dtype, x.device)
    return _vgg_crop_norm(torch.einsum('oh,bchw,pw->bcop', my, x, mx), top, left)


def vgg_features(vsd: Dict[str, Tensor], x: Tensor, taps=None) -> Tensor:
    """features[:36]: 16 x (conv3x3 + ReLU), max-pool after conv 2, 4, 8, 12 (utils/GAN.py:20-60, 69)."""
    idx = 0
    n = 0
    for c in VGG_CFG:
        if c == 'M':
            x = F.max_pool2d(x, 2, 2)
            idx += 1
        else:
            x = F.relu(F.conv2d(x, vsd[f'net.0.{idx}.weight'], vsd[f'net.0.{idx}.bias'], padding=1))
            if taps is not None:
                taps[f'v_y{n}'] = x
            n += 1
            idx += 2
    return x


def vgg_loss(vsd: Dict[str, Tensor], image1: Tensor, image2: Tensor) -> Tensor:
    """Vgg19Loss.forward (utils/GAN.py:74-88)."""
    return F.mse_loss(vgg_features(vsd, vgg_transform(image1)), vgg_features(vsd, vgg_transform(image2)))


def bce(p: Tensor, target: float) -> Tensor:
    return F.binary_cross_entropy(p, torch.full_like(p, target))


# ---------------------------------------------------------------------------------------------
# do_epoch (train_GAN.py:38-71) + Adam
# ---------------------------------------------------------------------------------------------
def adam_update(sd, grads, state, lr, t):
    """torch.optim.Adam defaults (betas 0.9 / 0.999, eps 1e-8, no weight decay), step t (1-based)."""
    b1, b2, eps = 0.9, 0.999, 1e-8
    for k, g in grads.items():
        m = state.setdefault('m.' + k, torch.zeros_like(g))
        v = state.setdefault('v.' + k, torch.zeros_like(g))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        step = lr / (1 - b1 ** t)
        denom = (v.sqrt() / (1 - b2 ** t) ** 0.5).add_(eps)
        sd[k] = sd[k] - step * (m / denom)


def _leaf(sd):
    return {k: (v.detach().clone().requires_grad_(True) if k in param_keys(sd) else v) for k, v in sd.items()}


def do_epoch(sdG, sdD, sdV, LR, HR, lr, factor=8, blocks=16, stG=None, stD=None, t=1):
    """One do_epoch.  Returns dict(loss_D, loss_G, gD (loss_D gradients), gG, fake); sdG / sdD are updated in place
    (parameters by Adam, running statistics by the three discriminator and two generator passes)."""
    stG = {} if stG is None else stG
    stD = {} if stD is None else stD
    out = {}
    # ---- discriminator step (:43-53)
    d = _leaf(sdD)
    ns = {}
    real = discriminator_train(d, HR, ns)
    for k, v in ns.items():
        d[k] = v
    with torch.no_grad():
        nsg = {}
        fake = generator_train(sdG, LR, factor, blocks, nsg)
        sdG.update(nsg)
    ns = {}
    fo = discriminator_train(d, fake, ns)
    loss_D = bce(real, 1.0) + bce(fo, 0.0)
    keys = param_keys(sdD)
    gD = dict(zip(keys, torch.autograd.grad(loss_D, [d[k] for k in keys])))
    sdD.update({k: v.detach() for k, v in ns.items()})
    adam_update(sdD, gD, stD, lr, t)
    # ---- generator step (:56-66)
    g = _leaf(sdG)
    nsg = {}
    fake2 = generator_train(g, LR, factor, blocks, nsg)
    ns = {}
    with torch.no_grad():
        fo2 = discriminator_train(sdD, fake2.detach(), ns)
        sdD.update(ns)
    content = vgg_loss(sdV, fake2, HR)
    loss_G = content + bce(fo2, 1.0)
    keys = param_keys(sdG)
    gG = dict(zip(keys, torch.autograd.grad(loss_G, [g[k] for k in keys], allow_unused=True)))
    gG = {k: (v if v is not None else torch.zeros_like(sdG[k])) for k, v in gG.items()}
    sdG.update({k: v.detach() for k, v in nsg.items()})
    adam_update(sdG, gG, stG, lr, t)
    out.update(loss_D=loss_D.detach(), loss_G=loss_G.detach(), content=content.detach(), gD=gD, gG=gG, fake=fake,
               p_real=real.detach(), p_fake=fo.detach(), p_fake2=fo2.detach())
    return out


def synthetic_batch(seed: int, batch: int, lr_hw: Tuple[int, int], factor: int = 8):
    """HR patches = smooth random fields in [0, 1]; LR = their area-averaged down-sampling (any fixed, deterministic
    LR / HR pair serves: the step's arithmetic does not depend on how the patches were made)."""
    g = torch.Generator().manual_seed(seed)
    h, w = lr_hw
    base = torch.rand(batch, 3, h, w, generator=g)
    hr = F.interpolate(base, scale_factor=factor, mode='bicubic', align_corners=False)
    hr = (hr + 0.05 * torch.randn(hr.shape, generator=g)).clamp(0, 1)
    lr = F.avg_pool2d(hr, factor)
    return lr, hr
