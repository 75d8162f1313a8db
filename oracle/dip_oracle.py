"""CPU oracle for the Deep-Image-Prior super-resolution step.  TEST INFRASTRUCTURE ONLY.

This file is the checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it; the product package (``deep-super-resolution_b200/``) never does and fails loudly when
its CUDA library is missing.

It is a *functional restatement* (plain fp32 tensor arithmetic, no ``nn.Module`` graph) of the
reference path, every function citing the reference ``file:line`` it follows (paths relative to
the upstream repo root).  The arithmetic primitives of the reference live in third-party PyTorch
(unpinned by the reference; 2.11.0+cu128 in this image), so the restatement is built on the same
``torch.nn.functional`` primitives.

Pinning: the reference ships no tests, fixtures or golden vectors ("parity unpinned" by the
reference itself).  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF, executed in the build container: ``oracle/make_golden.py`` imports the unmodified
reference modules from ``/root/reference`` and writes ``tests/golden/*.pt``; ``tests/test_oracle.py``
checks this restatement against those fixtures bit-for-bit (same torch, same thread count) or to
1e-6 otherwise.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BN_EPS = 1e-5          # torch.nn.BatchNorm2d default, used by models/DIP/utils.py:79-80
BN_MOMENTUM = 0.1
LRELU_SLOPE = 0.2      # models/DIP/utils.py:68


# --------------------------------------------------------------------------------------------
# Lanczos downsampler  (utils/downsampler.py)
# --------------------------------------------------------------------------------------------

def lanczos_taps_1d(factor: int, support: int = 2, phase: float = 0.5,
                    kernel_width: Optional[int] = None) -> np.ndarray:
    """Un-normalised 1-D Lanczos taps in float64.

    Follows utils/downsampler.py:103-127: the 2-D kernel entry (i, j) is the PRODUCT of a
    factor depending only on i and one depending only on j, so the kernel is rank-1 and these
    are its 1-D factors.  ``center=(kernel_width+1)/2`` (:105); for phase 0.5 the table has
    ``kernel_width-1`` entries (:77-78) and the distance is ``|i+0.5-center|/factor`` (:111).
    """
    if kernel_width is None:
        kernel_width = 4 * factor + 1 if support == 2 else 6 * factor + 1   # :14-22
    n = kernel_width - 1 if phase == 0.5 else kernel_width
    center = (kernel_width + 1) / 2.0
    taps = np.ones(n, dtype=np.float64)
    for i in range(1, n + 1):
        d = abs(i + 0.5 - center) / factor if phase == 0.5 else abs(i - center) / factor
        if d != 0:
            taps[i - 1] = support * np.sin(np.pi * d) * np.sin(np.pi * d / support) / (np.pi * np.pi * d * d)
    return taps


def lanczos_kernel_2d(factor: int, support: int = 2, phase: float = 0.5) -> np.ndarray:
    """Normalised 2-D kernel, float64 (utils/downsampler.py:73-134, 'lanczos' branch, :133 norm)."""
    t = lanczos_taps_1d(factor, support, phase)
    k = np.outer(t, t)
    return k / k.sum()


def downsampler_pad(factor: int, ksize: int) -> int:
    """preserve_size padding (utils/downsampler.py:53-61)."""
    if ksize % 2 == 1:
        return int((ksize - 1) / 2.0)
    return int((ksize - factor) / 2.0)


def downsample(x: Tensor, factor: int, kernel_type: str = 'lanczos2') -> Tensor:
    """Downsampler.forward (utils/downsampler.py:65-71) with phase 0.5, preserve_size=True:
    ReplicationPad2d(pad) then a per-plane (diagonal) strided correlation with the 2-D kernel
    cast to fp32 (:48-50), bias 0."""
    assert kernel_type in ('lanczos2', 'lanczos3')
    support = 2 if kernel_type == 'lanczos2' else 3
    k = torch.from_numpy(lanczos_kernel_2d(factor, support)).to(x.dtype).to(x.device)
    pad = downsampler_pad(factor, k.shape[0])
    c = x.shape[1]
    w = torch.zeros(c, c, *k.shape, dtype=x.dtype, device=x.device)
    for i in range(c):
        w[i, i] = k
    xp = F.pad(x, (pad, pad, pad, pad), mode='replicate')
    return F.conv2d(xp, w, bias=torch.zeros(c, dtype=x.dtype, device=x.device), stride=factor)


# --------------------------------------------------------------------------------------------
# Skip network (models/DIP/skip.py) -- state-dict naming
# --------------------------------------------------------------------------------------------

def level_prefix(i: int) -> str:
    """Children are numbered from 1 by Module.add (models/DIP/utils.py:5-8); level i+1 is the
    7th child of level i's `deeper` (skip.py:60-71), which is child '1' of Concat, itself child
    '1' of the level container."""
    return '1.1.7.' * i


def conv_child(pad: str = 'reflection') -> str:
    """Index of the Conv2d inside conv()'s Sequential (models/DIP/utils.py:96-105): 1 behind a ReflectionPad2d,
    0 with pad='zero' (no padder module)."""
    return '0' if pad == 'zero' else '1'


def layer_names(i: int, pad: str = 'reflection') -> Dict[str, str]:
    p = level_prefix(i)
    c = conv_child(pad)
    return dict(
        skip_conv=p + '1.0.1.' + c, skip_bn=p + '1.0.2',
        d1_conv=p + '1.1.1.' + c, d1_bn=p + '1.1.2',
        d2_conv=p + '1.1.4.' + c, d2_bn=p + '1.1.5',
        cat_bn=p + '2',
        u1_conv=p + '3.' + c, u1_bn=p + '4',
        u2_conv=p + '6.' + c, u2_bn=p + '7',
    )


FINAL_CONV = '9.1'   # skip.py:92: 9th child of the top container, conv() Sequential index 1 ('9.0' with pad='zero')


def init_params(input_depth: int = 32, n_out: int = 3, nd: int = 128, nu: int = 128, ns: int = 4,
                num_scales: int = 5, pad: str = 'reflection') -> Dict[str, Tensor]:
    """Fresh parameters + buffers with the reference's initialisation AND RNG consumption order.

    skip.py:41-92 constructs, per level, the skip 1x1 conv, the stride-2 conv, the second encoder
    conv, the decoder 3x3 conv and the decoder 1x1 conv (BatchNorm draws nothing), then the final
    1x1 conv; every ``nn.Conv2d`` draws weight (kaiming_uniform, a=sqrt(5)) then bias
    (U(+-1/sqrt(fan_in))).  Instantiating throw-away Conv2d's in that order reproduces
    ``get_net``'s same-seed weights bit-for-bit (checked in tests/test_oracle.py).
    """
    sd: Dict[str, Tensor] = {}

    def conv(name, cin, cout, k):
        m = torch.nn.Conv2d(cin, cout, k)
        sd[name + '.weight'] = m.weight.detach().clone()
        sd[name + '.bias'] = m.bias.detach().clone()

    def bn(name, c):
        sd[name + '.weight'] = torch.ones(c)
        sd[name + '.bias'] = torch.zeros(c)
        sd[name + '.running_mean'] = torch.zeros(c)
        sd[name + '.running_var'] = torch.ones(c)
        sd[name + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    cin = input_depth
    for i in range(num_scales):
        n = layer_names(i, pad)
        conv(n['skip_conv'], cin, ns, 1)
        conv(n['d1_conv'], cin, nd, 3)
        conv(n['d2_conv'], nd, nd, 3)
        conv(n['u1_conv'], ns + (nu if i < num_scales - 1 else nd), nu, 3)
        conv(n['u2_conv'], nu, nu, 1)
        cin = nd
    conv('9.' + conv_child(pad), nu, n_out, 1)
    # BN entries, in any order (no RNG); keys sorted later by the caller if needed
    for i in range(num_scales):
        n = layer_names(i, pad)
        bn(n['skip_bn'], ns)
        bn(n['d1_bn'], nd)
        bn(n['d2_bn'], nd)
        bn(n['cat_bn'], ns + (nu if i < num_scales - 1 else nd))
        bn(n['u1_bn'], nu)
        bn(n['u2_bn'], nu)
    return sd


def param_keys(sd: Dict[str, Tensor]) -> List[str]:
    """Learnable entries (what ``net.parameters()`` yields) in state-dict order."""
    return [k for k in sd if k.endswith('.weight') or k.endswith('.bias')]


def dead_param_keys(num_scales: int = 5, pad: str = 'reflection') -> List[str]:
    """Parameters whose value cannot influence the output: the bias of every conv that feeds a
    train-mode BatchNorm (cancelled by the mean subtraction) and BN(cat)'s beta (a per-channel
    constant on the input of reflect-pad + conv + train-mode BN).  The reference keeps and
    random-walks them (their gradients are rounding noise); weight-level parity skips them
    (SURVEY.md 7.2 item 3)."""
    out = []
    for i in range(num_scales):
        n = layer_names(i, pad)
        out += [n[k] + '.bias' for k in ('skip_conv', 'd1_conv', 'd2_conv', 'u1_conv', 'u2_conv')]
        if pad != 'zero':       # behind zero padding the constant is NOT uniform over the conv's input window
            out.append(n['cat_bn'] + '.bias')
    return out


# --------------------------------------------------------------------------------------------
# Skip network -- functional forward
# --------------------------------------------------------------------------------------------

Quant = Optional[Callable[[Tensor, str], Tensor]]


class _RoundSTE(torch.autograd.Function):
    """Round to a 16-bit format in the forward pass, identity in the backward pass."""

    @staticmethod
    def forward(ctx, t, dtype):
        return t.to(dtype).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


def fp16_points(dtype: torch.dtype = torch.float16) -> Callable[[Tensor, str], Tensor]:
    """Quantisation hook that rounds where the CUDA path stores 16-bit values (DESIGN.md "Numerics"): the operands of
    the 20 tensor-core convolutions ('conv_in', 'conv_w'), their raw outputs -- the BatchNorm inputs, whose statistics
    are taken from the rounded values ('raw') -- and every BatchNorm + LeakyReLU activation ('act': it feeds the
    bilinear upsample as well as the next convolution).  The skip-branch and final 1x1 convolutions keep fp32
    weights ('skip_w', 'final_w').  Gradients pass straight through the rounding (the CUDA backward pass
    differentiates the same rounded forward values)."""
    def q(t: Tensor, tag: str) -> Tensor:
        if tag in ('skip_w', 'final_w'):
            return t
        return _RoundSTE.apply(t, dtype)
    q.drop_bn_bias = True
    return q


def _q(q: Quant, t: Tensor, tag: str) -> Tensor:
    return t if q is None else q(t, tag)


def _conv(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int, q: Quant, w_tag: str = 'conv_w',
          pad: str = 'reflection') -> Tensor:
    """conv() of models/DIP/utils.py:83-105 with pad='reflection': ReflectionPad2d((k-1)/2)
    then Conv2d(padding=0).  ``w_tag`` names the weight operand for the quantisation hook: 'conv_w' for the 20
    tensor-core layers, 'skip_w' / 'final_w' for the small 1x1 layers (whose weights the CUDA path keeps in fp32)."""
    k = w.shape[-1]
    p = (k - 1) // 2
    if p:     # pad='zero': Conv2d(padding=p) (models/DIP/utils.py:96-102)
        x = F.pad(x, (p, p, p, p), mode='reflect') if pad != 'zero' else F.pad(x, (p, p, p, p))
    if q is not None and getattr(q, 'drop_bn_bias', False) and w_tag == 'conv_w':
        b = None      # a bias in front of a train-mode BatchNorm is cancelled by the mean subtraction; the CUDA kernels
                      # never add it, so their 16-bit rounding of the raw output happens WITHOUT it
    return F.conv2d(_q(q, x, 'conv_in'), _q(q, w, w_tag), b, stride=stride)


def _bn(x: Tensor, g: Tensor, b: Tensor) -> Tensor:
    """BatchNorm2d in train mode (models/DIP/utils.py:79-80; the net is never put in eval()):
    per-channel batch mean and BIASED variance over N*H*W, eps 1e-5.  Uses the same ATen primitive
    as nn.BatchNorm2d so that the restatement reproduces the reference's rounding."""
    return F.batch_norm(x, None, None, g, b, training=True, momentum=BN_MOMENTUM, eps=BN_EPS)


def _act(x: Tensor) -> Tensor:
    return F.leaky_relu(x, LRELU_SLOPE)          # models/DIP/utils.py:68


def _center_crop_cat(a: Tensor, b: Tensor) -> Tensor:
    """Concat.forward (models/DIP/utils.py:18-38): centre-crop every branch to the minimum
    H and W, then cat on dim 1."""
    h = min(a.shape[2], b.shape[2])
    w = min(a.shape[3], b.shape[3])
    outs = []
    for t in (a, b):
        d2 = (t.shape[2] - h) // 2
        d3 = (t.shape[3] - w) // 2
        outs.append(t[:, :, d2:d2 + h, d3:d3 + w])
    return torch.cat(outs, dim=1)


def skip_forward(sd: Dict[str, Tensor], z: Tensor, num_scales: int = 5, quant: Quant = None,
                 taps: Optional[Dict[str, Tensor]] = None, pad: str = 'reflection',
                 upsample_mode: str = 'bilinear') -> Tensor:
    """Forward of the net built by get_net(32,'skip','reflection',...,upsample_mode='bilinear')
    (models/DIP/skip.py:41-94; per-level structure in SURVEY.md 3.2).  ``taps`` (optional dict)
    receives named intermediates: 'L{i}.skip_raw', '.d1_raw', '.d2_raw', '.x_next', '.cat',
    '.u1_raw', '.u2_raw', '.out', and 'final_pre'."""

    def rec(i: int, x: Tensor) -> Tensor:
        n = layer_names(i, pad)
        P = lambda name: sd[name]
        # skip branch (skip.py:54-56)
        s_raw = _conv(x, P(n['skip_conv'] + '.weight'), P(n['skip_conv'] + '.bias'), 1, quant, 'skip_w', pad)
        s = _act(_bn(s_raw, P(n['skip_bn'] + '.weight'), P(n['skip_bn'] + '.bias')))
        # deeper branch (skip.py:60-66)
        d1_raw = _conv(x, P(n['d1_conv'] + '.weight'), P(n['d1_conv'] + '.bias'), 2, quant, pad=pad)
        d1 = _q(quant, _act(_bn(_q(quant, d1_raw, 'raw'), P(n['d1_bn'] + '.weight'), P(n['d1_bn'] + '.bias'))), 'act')
        d2_raw = _conv(d1, P(n['d2_conv'] + '.weight'), P(n['d2_conv'] + '.bias'), 1, quant, pad=pad)
        d2 = _q(quant, _act(_bn(_q(quant, d2_raw, 'raw'), P(n['d2_bn'] + '.weight'), P(n['d2_bn'] + '.bias'))), 'act')
        deep = rec(i + 1, d2) if i < num_scales - 1 else d2        # skip.py:70-75
        up = F.interpolate(deep, scale_factor=2, mode=upsample_mode)  # skip.py:77 (bilinear: align_corners False)
        cat = _center_crop_cat(s, up)                              # Concat(1, skip, deeper), skip.py:47
        c = _bn(cat, P(n['cat_bn'] + '.weight'), P(n['cat_bn'] + '.bias'))   # skip.py:51
        u1_raw = _conv(c, P(n['u1_conv'] + '.weight'), P(n['u1_conv'] + '.bias'), 1, quant, pad=pad)
        u1 = _q(quant, _act(_bn(_q(quant, u1_raw, 'raw'), P(n['u1_bn'] + '.weight'), P(n['u1_bn'] + '.bias'))), 'act')
        u2_raw = _conv(u1, P(n['u2_conv'] + '.weight'), P(n['u2_conv'] + '.bias'), 1, quant, pad=pad)
        u2 = _q(quant, _act(_bn(_q(quant, u2_raw, 'raw'), P(n['u2_bn'] + '.weight'), P(n['u2_bn'] + '.bias'))), 'act')
        if taps is not None:
            for k_, v_ in (('skip_raw', s_raw), ('d1_raw', d1_raw), ('d2_raw', d2_raw), ('x_next', d2),
                           ('up', up), ('cat', c), ('u1_raw', u1_raw), ('u2_raw', u2_raw), ('out', u2)):
                taps[f'L{i}.{k_}'] = v_
        return u2

    y = rec(0, z)
    fin = '9.' + conv_child(pad)
    pre = _conv(y, sd[fin + '.weight'], sd[fin + '.bias'], 1, quant, 'final_w', pad)   # skip.py:92
    if taps is not None:
        taps['final_pre'] = pre
    return torch.sigmoid(pre)                                                          # skip.py:93-94


def bn_running_update(sd: Dict[str, Tensor], name: str, x: Tensor) -> None:
    """Running-statistics side effect of a train-mode BatchNorm2d forward: momentum 0.1, the
    running variance uses the UNBIASED batch variance."""
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var_u = x.var(dim=(0, 2, 3), unbiased=True) if n > 1 else x.var(dim=(0, 2, 3), unbiased=False)
    sd[name + '.running_mean'] = (1 - BN_MOMENTUM) * sd[name + '.running_mean'] + BN_MOMENTUM * mean
    sd[name + '.running_var'] = (1 - BN_MOMENTUM) * sd[name + '.running_var'] + BN_MOMENTUM * var_u
    sd[name + '.num_batches_tracked'] = sd[name + '.num_batches_tracked'] + 1


# --------------------------------------------------------------------------------------------
# One optimisation step (closure of DIP.py:47-95 + utils/DIP.py:33-38)
# --------------------------------------------------------------------------------------------

def mse(a: Tensor, b: Tensor) -> Tensor:
    return ((a - b) ** 2).mean()                    # torch.nn.MSELoss() default 'mean', DIP.py:26


def step_loss_and_grads(sd: Dict[str, Tensor], z: Tensor, lr_image: Tensor, factor: int,
                        num_scales: int = 5, quant: Quant = None,
                        taps: Optional[Dict[str, Tensor]] = None, pad: str = 'reflection',
                        upsample_mode: str = 'bilinear'
                        ) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """net forward -> downsampler -> MSE -> backward (DIP.py:60-68).  Returns (loss, out_HR,
    grads by state-dict key).  ``z`` is the already perturbed input (DIP.py:52)."""
    keys = param_keys(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in keys}
    full = dict(sd)
    full.update(leaves)
    out = skip_forward(full, z, num_scales, quant, taps, pad, upsample_mode)
    if taps is not None:
        for t in taps.values():
            if t.requires_grad:
                t.retain_grad()
    loss = mse(downsample(out, factor), lr_image)
    loss.backward()
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])) for k in keys}
    return loss.detach(), out.detach(), grads


class AdamState:
    """torch.optim.Adam with its defaults as used by utils/DIP.py:34 (betas (0.9, 0.999),
    eps 1e-8, no weight decay, no amsgrad), single-tensor formulation:
        m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
        p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
    """

    def __init__(self, keys: Sequence[str], sd: Dict[str, Tensor], lr: float,
                 b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.t = 0
        self.m = {k: torch.zeros_like(sd[k]) for k in keys}
        self.v = {k: torch.zeros_like(sd[k]) for k in keys}

    def step(self, sd: Dict[str, Tensor], grads: Dict[str, Tensor]) -> None:
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        for k, g in grads.items():
            m, v = self.m[k], self.v[k]
            m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            sd[k] = sd[k] - (self.lr / bc1) * (m / denom)


def psnr(a: Tensor, b: Tensor) -> float:
    """10 log10(1/MSE) on [0,1] images (the explicit definition chosen in SURVEY.md 8c)."""
    return float(10.0 * torch.log10(1.0 / ((a - b) ** 2).mean()))


# --------------------------------------------------------------------------------------------
# Synthetic workload (SURVEY.md 8d / BASELINE.md 3)
# --------------------------------------------------------------------------------------------

def synthetic_hr(index: int, size: int) -> Tensor:
    """HR [3,H,H] in [0,1], deterministic per image index: bicubic upsample of U(0,1)
    [3,H/8,H/8] plus 0.3 x checker of period 32, clamped."""
    g = torch.Generator().manual_seed(1000 + index)
    low = torch.rand(1, 3, size // 8, size // 8, generator=g)
    field = F.interpolate(low, size=(size, size), mode='bicubic', align_corners=False)[0]
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    checker = (((yy // 16) + (xx // 16)) % 2).float() - 0.5
    return (field * 0.7 + 0.15 + 0.3 * checker).clamp(0, 1)


def synthetic_pair(index: int, size: int, factor: int = 4) -> Tuple[Tensor, Tensor]:
    hr = synthetic_hr(index, size)
    lr = downsample(hr.unsqueeze(0), factor)[0]
    return lr, hr
