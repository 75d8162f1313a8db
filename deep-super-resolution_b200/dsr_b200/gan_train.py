"""SRGAN training step: the call surface of train_GAN.py / utils/GAN.py / models/GAN/discriminator.py of the reference
over libdsr_b200.so (``dsr_gant_*``: bf16 tcgen05 convolutions, hand-written element-wise kernels, no cuDNN, no eager
fallback).

Two ways in, both backed by the same C entry points:

* drop-in modules -- ``Discriminator(HR_image_shape)`` (discriminator.py:21), ``Generator(factor)`` in ``.train()``
  mode (generator.py:44), ``Vgg19Loss`` / ``PerceptualLoss`` / ``get_loss_D`` / ``get_adversarial_loss``
  (utils/GAN.py:6-123).  Their ``forward`` is a ``torch.autograd.Function`` around one library call, their parameters
  are views of one flat fp32 buffer per network, so the reference's own ``do_epoch`` (train_GAN.py:38-71) with its
  ``loss.backward()`` / ``torch.optim.Adam`` runs over them unchanged;
* ``GanTrainStep.do_epoch(LR, HR)`` -- the same step as one fused sequence (one generator forward instead of two
  identical ones, BCE fused into the discriminator backward, fused flat Adam) and, under ``torch.distributed``, the
  data-parallel version: per-replica BatchNorm statistics, NCCL all-reduce (mean) of the flat discriminator gradient
  (80 M floats) overlapped with the generator phase, then of the generator gradient.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check

NET_G, NET_D, NET_VGG = 0, 1, 2
_VGG_CONV_IDX = (0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28, 30, 32, 34)


# =================================================================================================
# trainer object (one per batch / patch size / device): workspace + C handle
# =================================================================================================
class Trainer:
    GUARD = 4096

    def __init__(self, batch: int, lr_h: int, lr_w: int, factor: int, blocks: int, device: torch.device,
                 with_vgg: bool = True):
        _lib.require_cuda()
        self.key = (batch, lr_h, lr_w, factor, blocks)
        self.B, self.lr_h, self.lr_w, self.factor, self.blocks = batch, lr_h, lr_w, factor, blocks
        self.H, self.W = lr_h * factor, lr_w * factor
        self.device = device
        self.handle = C.c_void_p()
        check(lib.dsr_gant_create(C.byref(self.handle), batch, lr_h, lr_w, factor, blocks, int(with_vgg)), 'dsr_gant_create')
        nbytes = lib.dsr_gant_workspace_bytes(self.handle)
        with torch.cuda.device(device):
            self.workspace = torch.empty(nbytes + 1024 + self.GUARD, dtype=torch.uint8, device=device)
            base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
            self._guard = self.workspace[base - self.workspace.data_ptr() + nbytes:][:self.GUARD]
            self._guard.fill_(0xA5)          # canary behind the workspace (guard_intact(), tests)
            check(lib.dsr_gant_bind(self.handle, base, nbytes, _lib.stream_ptr()), 'dsr_gant_bind')
        self.d_slot = 0
        self.d_gen = [0, 0]          # generation counter per discriminator activation slot
        self.g_gen = 0
        self._packed: Dict[int, Tuple[int, int]] = {}
        self.launch_count = 0        # kernels launched through this trainer (dsr_gant_last_launches per call)

    def _count(self) -> None:
        self.launch_count += max(0, lib.dsr_gant_last_launches(self.handle))

    def layout(self, net: int):
        name = C.create_string_buffer(160)
        off, n = C.c_longlong(), C.c_longlong()
        params, bufs = [], []
        for i in range(lib.dsr_gant_num_params(self.handle, net)):
            check(lib.dsr_gant_param_info(self.handle, net, i, name, 160, C.byref(off), C.byref(n)))
            params.append((name.value.decode(), off.value, n.value))
        for i in range(lib.dsr_gant_num_buffers(self.handle, net)):
            check(lib.dsr_gant_buffer_info(self.handle, net, i, name, 160, C.byref(off), C.byref(n)))
            bufs.append((name.value.decode(), off.value, n.value))
        return dict(params=params, buffers=bufs, nparam=lib.dsr_gant_param_numel(self.handle, net),
                    nbuf=lib.dsr_gant_buffer_numel(self.handle, net))

    # ---- thin wrappers -------------------------------------------------------------------------
    def pack(self, net: int, flat: torch.Tensor, force: bool = False) -> None:
        key = (flat.data_ptr(), flat._version)
        if not force and self._packed.get(net) == key:
            return
        check(lib.dsr_gant_pack(self.handle, net, flat.data_ptr(), _lib.stream_ptr()), 'dsr_gant_pack')
        self._packed[net] = key
        self._count()

    def g_forward(self, params, buffers, lr, bn_updates: int = 1) -> torch.Tensor:
        out = torch.empty((self.B, 3, self.H, self.W), dtype=torch.float32, device=lr.device)
        check(lib.dsr_gant_g_forward(self.handle, params.data_ptr(), buffers.data_ptr() if buffers is not None else None,
                                     lr.data_ptr(), out.data_ptr(), bn_updates, _lib.stream_ptr()), 'dsr_gant_g_forward')
        self.g_gen += 1
        self._count()
        return out

    def g_backward(self, params, dout, grads) -> None:
        check(lib.dsr_gant_g_backward(self.handle, params.data_ptr(), dout.data_ptr(), grads.data_ptr(), _lib.stream_ptr()),
              'dsr_gant_g_backward')
        self._count()

    def d_forward(self, slot: int, params, buffers, img) -> torch.Tensor:
        prob = torch.empty((self.B,), dtype=torch.float32, device=img.device)
        check(lib.dsr_gant_d_forward(self.handle, slot, params.data_ptr(),
                                     buffers.data_ptr() if buffers is not None else None, img.data_ptr(), prob.data_ptr(),
                                     _lib.stream_ptr()), 'dsr_gant_d_forward')
        self.d_gen[slot] += 1
        self._count()
        return prob

    def d_backward(self, slot: int, params, grads, dprob: Optional[torch.Tensor] = None, target: float = 0.0) -> None:
        check(lib.dsr_gant_d_backward(self.handle, slot, params.data_ptr(), dprob.data_ptr() if dprob is not None else None,
                                      float(target), grads.data_ptr(), _lib.stream_ptr()), 'dsr_gant_d_backward')
        self._count()

    def d_backward_pair(self, params, grads, target0: float, target1: float, dense_overwrite: bool = False) -> None:
        """``dense_overwrite``: the gradient of dense1.weight is written, not added to (the caller skipped clearing it)."""
        if dense_overwrite:
            check(lib.dsr_gant_dense_grad_overwrite(self.handle, 1))
        try:
            check(lib.dsr_gant_d_backward_pair(self.handle, params.data_ptr(), float(target0), float(target1),
                                               grads.data_ptr(), _lib.stream_ptr()), 'dsr_gant_d_backward_pair')
        finally:
            if dense_overwrite:
                check(lib.dsr_gant_dense_grad_overwrite(self.handle, 0))
        self._count()

    def bce(self, prob, target: float, loss, accumulate: bool) -> None:
        check(lib.dsr_gant_bce(self.handle, prob.data_ptr(), float(target), prob.numel(), loss.data_ptr(), int(accumulate),
                               _lib.stream_ptr()), 'dsr_gant_bce')
        self.launch_count += 1

    def vgg_real(self, real) -> None:
        """VGG features of the real batch alone; ``vgg_loss(fake, None, ...)`` then uses them."""
        check(lib.dsr_gant_vgg_real(self.handle, real.data_ptr(), _lib.stream_ptr()), 'dsr_gant_vgg_real')
        self._count()

    def vgg_loss(self, fake, real, loss, accumulate: bool, want_grad: bool) -> Optional[torch.Tensor]:
        dfake = torch.empty_like(fake) if want_grad else None
        check(lib.dsr_gant_vgg_loss(self.handle, fake.data_ptr(), real.data_ptr() if real is not None else None,
                                    loss.data_ptr(), int(accumulate),
                                    dfake.data_ptr() if want_grad else None, _lib.stream_ptr()), 'dsr_gant_vgg_loss')
        self._count()
        return dfake

    def device_error(self) -> int:
        code = C.c_int()
        check(lib.dsr_gant_device_error(self.handle, C.byref(code)))
        return code.value

    def guard_intact(self) -> bool:
        return bool((self._guard == 0xA5).all())

    def tensor(self, name: str, with_gap: bool = False) -> torch.Tensor:
        """Named activation of the last pass as [B, H, W, C] float32 (tests); with_gap: [B, P, W, C] incl. the gap rows."""
        ptr = C.c_void_p()
        cc, w, h, p, b, f32 = (C.c_int() for _ in range(6))
        check(lib.dsr_gant_tensor(self.handle, name.encode(), C.byref(ptr), C.byref(cc), C.byref(w), C.byref(h), C.byref(p),
                                  C.byref(b), C.byref(f32)), f'dsr_gant_tensor({name})')
        dt = {0: torch.bfloat16, 1: torch.float32, 2: torch.float16}[f32.value]
        t = torch.empty((b.value * p.value, w.value, cc.value), dtype=dt, device=self.device)
        check(lib.dsr_debug_copy(t.data_ptr(), ptr, t.numel() * t.element_size(), _lib.stream_ptr()), 'dsr_debug_copy')
        t = t.view(b.value, p.value, w.value, cc.value)
        return t.float() if with_gap else t[:, :h.value].float()

    def __del__(self):
        try:
            if self.handle:
                lib.dsr_gant_destroy(self.handle)
        except Exception:
            pass


_TRAINERS: Dict[Tuple, Trainer] = {}


def get_trainer(batch: int, H: int, W: int, device: torch.device, factor: Optional[int] = None, blocks: int = 16) -> Trainer:
    """The trainer whose HR patch size is (H, W); the generator states its factor, the discriminator / VGG side take
    whichever trainer of that HR size exists (or a factor-8 one)."""
    dev = device.index if device.index is not None else torch.cuda.current_device()
    for (b, h, w, d, f, k), t in _TRAINERS.items():
        if (b, h, w, d) == (batch, H, W, dev) and (factor is None or (f == factor and k == blocks)):
            return t
    f = factor or 8
    if H % f or W % f:
        raise ValueError(f'patch size {H} x {W} is no multiple of the factor {f}')
    t = Trainer(batch, H // f, W // f, f, blocks, torch.device('cuda', dev))
    _TRAINERS[(batch, H, W, dev, f, blocks)] = t
    return t


# =================================================================================================
# flat parameter storage shared by the drop-in modules
# =================================================================================================
class FlatParams:
    """Re-points a module's parameters (and BatchNorm running statistics) at views of flat fp32 device buffers laid out
    as the library expects; gradients are views of a second flat buffer."""

    def __init__(self, module: nn.Module, net: int, key_prefix: str = ''):
        self.module = weakref.ref(module)
        self.net = net
        self.flat = self.gflat = self.bflat = None
        self.grad_views: List[torch.Tensor] = []
        self.grads_valid = False
        self.layout = None

    def ensure(self, trainer: Trainer, device: torch.device) -> None:
        m = self.module()
        named = list(m.named_parameters())
        if self.flat is not None and self.flat.device == device and named[0][1].data_ptr() == self.flat.data_ptr():
            return
        lay = trainer.layout(self.net)
        if [n for n, _ in named] != [n for n, _, _ in lay['params']]:
            raise RuntimeError('dsr_b200: parameter order of the module and of the library disagree')
        flat = torch.empty(lay['nparam'], dtype=torch.float32, device=device)
        gflat = torch.zeros(lay['nparam'], dtype=torch.float32, device=device)
        views = []
        with torch.no_grad():
            for (name, p), (_, off, n) in zip(named, lay['params']):
                if p.numel() != n:
                    raise RuntimeError(f'dsr_b200: {name} has {p.numel()} elements, the library expects {n}')
                v = flat[off:off + n].view(p.shape)
                v.copy_(p.data)
                p.data = v
                p.grad = None
                views.append(gflat[off:off + n].view(p.shape))
            bflat = torch.empty(max(lay['nbuf'], 1), dtype=torch.float32, device=device)
            bufs = dict(m.named_buffers())
            for name, off, n in lay['buffers']:
                v = bflat[off:off + n]
                v.copy_(bufs[name])
                mod_name, _, leaf = name.rpartition('.')
                setattr(m.get_submodule(mod_name), leaf, v)
        self.flat, self.gflat, self.bflat, self.grad_views, self.layout = flat, gflat, bflat, views, lay
        self.grads_valid = False

    def begin_backward(self) -> None:
        if not self.grads_valid:
            self.gflat.zero_()
            self.grads_valid = True

    def publish_grads(self) -> None:
        for (_, p), v in zip(self.module().named_parameters(), self.grad_views):
            if p.requires_grad and p.grad is None:
                p.grad = v

    def invalidate(self) -> None:
        self.grads_valid = False


def _bump_batches_tracked(module: nn.Module, times: int = 1) -> None:
    for m in module.modules():
        if isinstance(m, nn.BatchNorm2d) and m.num_batches_tracked is not None:
            m.num_batches_tracked += times


# =================================================================================================
# Discriminator (models/GAN/discriminator.py)
# =================================================================================================
class DiscriminatorConvBlock(nn.Module):
    """Parameter container with the keys of discriminator.py:4-12 (conv1, bn1)."""

    def __init__(self, in_channels: int, out_channels: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride, 1)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.leakyrelu = nn.LeakyReLU(0.2)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError('dsr_b200: blocks are parameter containers; call the Discriminator')


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, anchor, module):
        tr, fp = module._trainer_for(img)
        slot = tr.d_slot
        tr.d_slot ^= 1
        tr.pack(NET_D, fp.flat, force=True)
        prob = tr.d_forward(slot, fp.flat, fp.bflat, img)
        _bump_batches_tracked(module)
        ctx.module, ctx.trainer, ctx.slot, ctx.gen = module, tr, slot, tr.d_gen[slot]
        return prob.view(-1, 1)

    @staticmethod
    def backward(ctx, dprob):
        tr, module = ctx.trainer, ctx.module
        if tr.d_gen[ctx.slot] != ctx.gen:
            raise RuntimeError('dsr_b200.Discriminator: the activations of this forward pass were overwritten by two later '
                               'passes before backward() (two passes are kept, as do_epoch needs)')
        fp = module._fp
        fp.begin_backward()
        tr.d_backward(ctx.slot, fp.flat, fp.gflat, dprob=dprob.contiguous().view(-1).float())
        fp.publish_grads()
        return None, None, None


class Discriminator(nn.Module):
    """``Discriminator(HR_image_shape)`` of discriminator.py:21-74: same module tree, names, construction order (hence
    state_dict keys and same-seed initialisation).  Train mode only (train_GAN.py:164 never leaves it)."""

    def __init__(self, HR_image_shape):
        super().__init__()
        self.conv = nn.Conv2d(3, 64, 3, 1, 1)
        self.leakyrelu1 = nn.LeakyReLU(0.2)
        self.convblocks = nn.Sequential(*[DiscriminatorConvBlock(a, b, s) for a, b, s in
                                          ((64, 64, 2), (64, 128, 1), (128, 128, 2), (128, 256, 1), (256, 256, 2),
                                           (256, 512, 1), (512, 512, 2))])
        self.dense1 = nn.Linear(self.fc_input_shape(HR_image_shape), 1024)
        self.leakyrelu2 = nn.LeakyReLU(0.2)
        self.dense2 = nn.Linear(1024, 1)
        self.sigmoid = nn.Sigmoid()
        self.hr_shape = (int(HR_image_shape[0]), int(HR_image_shape[1]))
        self._fp = FlatParams(self, NET_D)

    def fc_input_shape(self, HR_image_shape):
        """discriminator.py:47-55, construction time only: a batch of ones goes through conv / convblocks in TRAIN mode
        (so the BatchNorm running statistics take one update before training starts -- kept, it is part of the
        reference's same-seed initial state) to read off dense1's input width.  Stock torch ops on the CPU, once."""
        with torch.no_grad():
            x = torch.ones(1, 3, int(HR_image_shape[0]), int(HR_image_shape[1]))
            x = self.leakyrelu1(self.conv(x))
            for blk in self.convblocks:
                x = blk.leakyrelu(blk.bn1(blk.conv1(x)))
            return x.view(x.size(0), -1).shape[1]

    def _trainer_for(self, img):
        B, _, H, W = img.shape
        tr = get_trainer(B, H, W, img.device)
        self._fp.ensure(tr, img.device)
        return tr, self._fp

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._fp.invalidate()
        for p in self.parameters():
            p.grad = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError('dsr_b200.Discriminator needs a CUDA tensor (sm_100a); there is no CPU fallback')
        if not self.training:
            raise NotImplementedError('dsr_b200.Discriminator: eval mode is not built (train_GAN.py keeps it in train mode)')
        if x.dim() != 4 or x.shape[1] != 3 or tuple(x.shape[2:]) != self.hr_shape:
            raise ValueError(f'expected [B, 3, {self.hr_shape[0]}, {self.hr_shape[1]}], got {tuple(x.shape)}')
        if x.requires_grad:
            raise NotImplementedError('dsr_b200.Discriminator: the gradient w.r.t. the input image is not built -- do_epoch '
                                      'detaches the generated image in front of the discriminator (train_GAN.py:46,58)')
        x = x.detach().to(torch.float32).contiguous()
        return _DiscFn.apply(x, next(self.parameters()), self)


# =================================================================================================
# Generator in train mode (the module itself lives in gan.py; this is its training forward)
# =================================================================================================
class _GenTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lr, anchor, module):
        tr, fp = module._train_state(lr)
        tr.pack(NET_G, fp.flat, force=True)
        out = tr.g_forward(fp.flat, fp.bflat, lr, 1)
        _bump_batches_tracked(module)
        ctx.module, ctx.trainer, ctx.gen = module, tr, tr.g_gen
        return out

    @staticmethod
    def backward(ctx, dout):
        tr, module = ctx.trainer, ctx.module
        if tr.g_gen != ctx.gen:
            raise RuntimeError('dsr_b200.Generator: a later forward pass overwrote the activations of this one before '
                               'backward()')
        fp = module._fp
        fp.begin_backward()
        tr.g_backward(fp.flat, dout.contiguous().float(), fp.gflat)
        fp.publish_grads()
        return None, None, None


def generator_train_forward(module, x: torch.Tensor) -> torch.Tensor:
    if x.requires_grad:
        raise NotImplementedError('dsr_b200.Generator: the gradient w.r.t. the LR input is not built')
    x = x.detach().to(torch.float32).contiguous()
    return _GenTrainFn.apply(x, module.conv1.weight, module)


# =================================================================================================
# losses (utils/GAN.py)
# =================================================================================================
class _VggLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fake, real, module):
        tr = module._trainer_for(fake)
        loss = torch.empty((), dtype=torch.float32, device=fake.device)
        want = fake.requires_grad
        dfake = tr.vgg_loss(fake.detach().contiguous().float(), real.detach().contiguous().float(), loss, False, want)
        ctx.dfake = dfake
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.dfake * g if ctx.dfake is not None else None), None, None


class Vgg19Loss(nn.Module):
    """utils/GAN.py:6-88.  ``self.net`` holds torchvision's ``vgg19().features[:36]`` as a frozen parameter container
    (keys ``net.0.<idx>.weight``); the forward pass -- transform, 16 convolutions, 4 max-pools, feature MSE and the
    gradient w.r.t. the first image -- is one library call.  The reference downloads the IMAGENET1K_V1 weights; with no
    cached copy (no network here) the container keeps torchvision's default random initialisation unless a state dict
    is loaded into it."""

    def __init__(self, pretrained: bool = True):
        super().__init__()
        from torchvision.models import vgg19
        layers = None
        import os
        if pretrained and not os.environ.get('DSR_VGG_RANDOM'):
            try:
                from torchvision.models import VGG19_Weights
                layers = vgg19(weights=VGG19_Weights.IMAGENET1K_V1).features
            except Exception:                           # no cache and no network: random weights
                layers = None
        if layers is None:
            layers = vgg19(weights=None).features
        self.net = nn.Sequential(layers[:36])
        self.mse = nn.MSELoss()
        for p in self.net.parameters():
            p.requires_grad = False
        self._flat = None

    def pack_into(self, tr: Trainer, device) -> None:
        """Flat fp32 copy of the (frozen) weights in library order, packed into the trainer's GEMM layouts; redone only
        when a weight tensor changed (load_state_dict)."""
        key = tuple((p.data_ptr(), p._version) for p in self.net.parameters())
        if self._flat is None or self._flat[0] != key or self._flat[1].device != torch.device(device):
            flat = torch.cat([p.detach().to(device=device, dtype=torch.float32).reshape(-1)
                              for p in self.net.parameters()]).contiguous()
            assert flat.numel() == lib.dsr_gant_param_numel(tr.handle, NET_VGG)
            self._flat = (key, flat)
        tr.pack(NET_VGG, self._flat[1])

    def _trainer_for(self, img):
        B, _, H, W = img.shape
        tr = get_trainer(B, H, W, img.device)
        self.pack_into(tr, img.device)
        return tr

    def forward(self, image1, image2):
        if not image1.is_cuda:
            raise RuntimeError('dsr_b200.Vgg19Loss needs CUDA tensors (sm_100a); there is no CPU fallback')
        return _VggLossFn.apply(image1, image2, self)


def get_adversarial_loss(fake_output, bce_loss):            # utils/GAN.py:96-98
    return bce_loss(fake_output, torch.ones_like(fake_output))


def get_loss_D(real_output, fake_output, bce_loss):         # utils/GAN.py:101-107
    real_loss = bce_loss(real_output, torch.ones_like(real_output))
    fake_loss = bce_loss(fake_output, torch.zeros_like(fake_output))
    return real_loss + fake_loss


class PerceptualLoss(nn.Module):                            # utils/GAN.py:110-123
    def __init__(self, pretrained: bool = True):
        super().__init__()
        self.vgg_loss = Vgg19Loss(pretrained)

    def forward(self, fake_output_G, HR_images, fake_output_D, bce_loss):
        content_loss = self.vgg_loss(fake_output_G, HR_images)
        adversarial_loss_ = get_adversarial_loss(fake_output_D, bce_loss)
        return content_loss + adversarial_loss_


# =================================================================================================
# data-parallel gradient exchange (the one collective of the repository's data paths)
# =================================================================================================
class GradExchange:
    """Mean all-reduce of flat gradient buffers over the default process group (NCCL over NVLink / NVSwitch on a GPU
    box, gloo in the CPU tests) and the initial broadcast of rank 0's state.  ``start`` enqueues the collective -- for
    CUDA tensors on a side stream that first waits for the producer stream, so that it overlaps whatever the caller
    launches next; ``finish`` makes the caller's stream wait for it and divides by the world size (DDP's averaging)."""

    def __init__(self, device=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('GradExchange needs an initialised torch.distributed process group')
        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.device = torch.device(device) if device is not None else None
        self.stream = torch.cuda.Stream(device=self.device) if (self.device is not None and self.device.type == 'cuda') else None
        self.bytes = 0

    def broadcast(self, *tensors: torch.Tensor, src: int = 0) -> None:
        for t in tensors:
            self.dist.broadcast(t, src)

    def start(self, flat: torch.Tensor):
        self.bytes += flat.numel() * flat.element_size()
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                return self.dist.all_reduce(flat, async_op=True)
        return self.dist.all_reduce(flat, async_op=True)

    def finish(self, work, flat: torch.Tensor) -> None:
        work.wait()
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        flat.mul_(1.0 / self.world)

    def allreduce_mean(self, flat: torch.Tensor) -> None:
        self.finish(self.start(flat), flat)


# =================================================================================================
# fused / data-parallel step
# =================================================================================================
class GanTrainStep:
    """``do_epoch`` of train_GAN.py:38-71 as one fused sequence over flat buffers, optionally data-parallel.

    ``gan_G`` / ``gan_D`` are the drop-in modules (their parameters become views of the flat buffers this object
    updates), ``vgg`` a ``Vgg19Loss``.  With an initialised ``torch.distributed`` process group (``data_parallel=True``)
    every rank holds a replica, rank 0's initial parameters are broadcast, each rank processes its own batch shard with
    per-replica BatchNorm statistics (what DistributedDataParallel does), and the flat gradients are averaged with NCCL
    all-reduce: the discriminator's 321 MB on a side stream while the generator phase (VGG + generator backward) runs,
    the generator's 6 MB after it."""

    def __init__(self, gan_G, gan_D, vgg: Vgg19Loss, lr: float, batch: int, lr_hw: Tuple[int, int], device,
                 data_parallel: bool = False):
        device = torch.device(device)
        self.G, self.D, self.vgg, self.lr = gan_G, gan_D, vgg, float(lr)
        self.tr = get_trainer(batch, lr_hw[0] * gan_G.factor, lr_hw[1] * gan_G.factor, device, gan_G.factor, gan_G.blocks)
        with torch.cuda.device(device):
            gan_G.to(device); gan_D.to(device)
            gan_G._fp.ensure(self.tr, device)
            gan_D._fp.ensure(self.tr, device)
            self.fg, self.fd = gan_G._fp, gan_D._fp
            self.mG, self.vG = torch.zeros_like(self.fg.flat), torch.zeros_like(self.fg.flat)
            self.mD, self.vD = torch.zeros_like(self.fd.flat), torch.zeros_like(self.fd.flat)
            self.loss_D = torch.zeros((), device=device)
            self.loss_G = torch.zeros((), device=device)
        self.t = 0
        self.dp = bool(data_parallel)
        self.world = 1
        self.xch = None
        if self.dp:
            self.xch = GradExchange(device)
            self.world = self.xch.world
            self.xch.broadcast(self.fg.flat, self.fg.bflat, self.fd.flat, self.fd.bflat)
        self.device = device
        self._pending_batches = [0, 0]
        self._need_pack = True           # set it again after changing parameters outside do_epoch (load_state_dict)
        span = [(off, off + n) for name, off, n in self.tr.layout(NET_D)['params'] if name == 'dense1.weight']
        self._dense_span = span[0]
        self._one_stream = bool(os.environ.get('DSR_GAN_ONE_STREAM'))
        self._side = None if self._one_stream else torch.cuda.Stream(device=device)
        self._side2 = (torch.cuda.Stream(device=device)
                       if (not self._one_stream and not os.environ.get('DSR_GAN_VGG_REAL_LATE')) else None)
        with torch.cuda.device(device):
            vgg.pack_into(self.tr, device)

    def flush_counters(self) -> None:
        """num_batches_tracked of the BatchNorm modules (2 generator / 3 discriminator passes per step) is kept on the
        host during the fused loop and written to the modules here (before state_dict() / saving)."""
        _bump_batches_tracked(self.G, self._pending_batches[0])
        _bump_batches_tracked(self.D, self._pending_batches[1])
        self._pending_batches = [0, 0]

    def _adam(self, p, g, m, v) -> None:
        check(lib.dsr_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), self.lr, 0.9, 0.999,
                                1e-8, self.t, _lib.stream_ptr()), 'dsr_adam_step')
        self.tr.launch_count += 1

    def do_epoch(self, LR_patches: torch.Tensor, HR_patches: torch.Tensor):
        tr, fg, fd = self.tr, self.fg, self.fd
        LR = LR_patches.to(self.device, non_blocking=True).float().contiguous()
        HR = HR_patches.to(self.device, non_blocking=True).float().contiguous()
        self.t += 1
        n0 = tr.launch_count
        # The step is TWO chains that only share `fake` (train_GAN.py:43-66: the discriminator sees the generated batch
        # detached, the generator's gradient is the content loss's alone):
        #   D chain  D(HR) | D(fake), BCE, backward of both passes, [all-reduce], Adam(D), D(fake) again, BCE -> loss_G
        #   G chain  G(LR) | VGG loss + gradient, generator backward, [all-reduce], Adam(G)
        # Most of their launches run at 24 x 24 .. 48 x 48 where a launch costs its latency, not its work, so the D chain
        # goes to a second stream and the two overlap (DSR_GAN_ONE_STREAM=1: everything on the caller's stream); the VGG
        # features of the REAL batch run on a third one beside the generator's forward pass (6.69 -> 6.55 ms; a high-
        # priority stream for the G chain made it slower: 6.88 ms).
        s0 = torch.cuda.current_stream(self.device)
        s1 = s0 if self._one_stream else self._side
        if self._need_pack:
            tr.pack(NET_D, fd.flat, force=True)
            tr.pack(NET_G, fg.flat, force=True)
            self._need_pack = False
        s1.wait_stream(s0)
        if self._side2 is not None:                  # VGG features of the real batch while the generator runs
            self._side2.wait_stream(s0)
            with torch.cuda.stream(self._side2):
                tr.vgg_real(HR)
        with torch.cuda.stream(s1):
            p_real = tr.d_forward(0, fd.flat, fd.bflat, HR)
        fake = tr.g_forward(fg.flat, fg.bflat, LR, bn_updates=2)        # the two generator passes of do_epoch are identical
        s1.wait_stream(s0)                                              # `fake` is complete (before s0 joins the VGG pass)
        if self._side2 is not None:
            s0.wait_stream(self._side2)
        with torch.cuda.stream(s1):
            # ---- discriminator step (train_GAN.py:43-53), its update, and its pass on the generated batch for the
            # adversarial term of loss_G (:58-59).  The bf16 GEMM copies of the weights are refreshed after each Adam step
            p_fake = tr.d_forward(1, fd.flat, fd.bflat, fake)
            tr.bce(p_real, 1.0, self.loss_D, False)
            tr.bce(p_fake, 0.0, self.loss_D, True)
            # zero_grad() + loss_D.backward(): dense1.weight's gradient (302 of the 321 MB) is written by the backward
            # pass instead of cleared here and added to there
            a, b = self._dense_span
            fd.gflat[:a].zero_()
            fd.gflat[b:].zero_()
            tr.d_backward_pair(fd.flat, fd.gflat, 1.0, 0.0, dense_overwrite=True)
            if self.dp:
                self.xch.allreduce_mean(fd.gflat)
            self._adam(fd.flat, fd.gflat, self.mD, self.vD)
            tr.pack(NET_D, fd.flat, force=True)
            p_fake2 = tr.d_forward(0, fd.flat, fd.bflat, fake)
        # ---- generator phase: content loss and generator backward (:56-66)
        dfake = tr.vgg_loss(fake, None if self._side2 is not None else HR, self.loss_G, False, True)
        s1.wait_stream(s0)                                              # loss_G holds the content term before BCE is added
        with torch.cuda.stream(s1):
            tr.bce(p_fake2, 1.0, self.loss_G, True)
        fg.gflat.zero_()
        tr.g_backward(fg.flat, dfake, fg.gflat)
        if self.dp:
            self.xch.allreduce_mean(fg.gflat)
        self._adam(fg.flat, fg.gflat, self.mG, self.vG)
        tr.pack(NET_G, fg.flat, force=True)
        s0.wait_stream(s1)
        self.launches_per_step = tr.launch_count - n0
        self._pending_batches[0] += 2
        self._pending_batches[1] += 3
        return self.loss_D, self.loss_G
