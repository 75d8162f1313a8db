"""SRResNet generator, inference: `Generator` with the reference's call surface.

Mirrors models/GAN/generator.py:4-81 of the reference as eval_GAN.py:87-94 uses it: ``Generator(factor=factor)
.to(device)``, ``load_state_dict`` (utils/common.py:46-59), ``.eval()``, ``gan_G(LR_image)``.  The module tree (names,
shapes, construction order, hence ``state_dict()`` keys and same-seed initialisation) is the reference's; the
sub-modules are parameter containers only -- ``forward`` is one call into libdsr_b200.so (``dsr_gen_forward``: tcgen05
convolutions with eval-mode BatchNorm folded in, PReLU / residual adds / PixelShuffle fused into the conv epilogues).

In ``.train()`` mode (train_GAN.py) ``forward`` is the training pass of dsr_b200/gan_train.py: batch-statistics
BatchNorm, activations kept for ``backward()`` (``dsr_gant_g_forward`` / ``dsr_gant_g_backward``).  CPU tensors raise,
there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check

_SHUFFLES = {8: 3, 16: 4}          # generator.py:55-58


class ResidualBlock(nn.Module):
    """Parameter container with the keys of generator.py:4-12 (conv1, bn1, prelu1, conv2, bn2)."""

    def __init__(self):
        super().__init__()
        for i in (1, 2):
            setattr(self, f'conv{i}', nn.Conv2d(64, 64, 3, 1, 1))
            setattr(self, f'bn{i}', nn.BatchNorm2d(64))
            if i == 1:
                self.prelu1 = nn.PReLU()

    def forward(self, x):  # pragma: no cover
        raise RuntimeError('dsr_b200: blocks are parameter containers; call the Generator')


class PixelShuffleBlock(nn.Module):
    """Parameter container with the keys of generator.py:27-34 (conv1 64 -> 256, prelu1)."""

    def __init__(self, in_channels: int = 64):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 256, 3, 1, 1)
        self.prelu1 = nn.PReLU()

    def forward(self, x):  # pragma: no cover
        raise RuntimeError('dsr_b200: blocks are parameter containers; call the Generator')


class _GenPlan:
    def __init__(self, factor: int, blocks: int, batch: int, h: int, w: int, device: torch.device):
        self.handle = C.c_void_p()
        check(lib.dsr_gen_plan_create(C.byref(self.handle), factor, blocks, batch, h, w), 'dsr_gen_plan_create')
        nbytes = lib.dsr_gen_workspace_bytes(self.handle)
        self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        check(lib.dsr_gen_bind(self.handle, base, nbytes, _lib.stream_ptr()), 'dsr_gen_bind')
        self.loaded_key = None

    def names(self):
        out = []
        buf = C.create_string_buffer(128)
        off, n = C.c_longlong(), C.c_longlong()
        for i in range(lib.dsr_gen_num_tensors(self.handle)):
            check(lib.dsr_gen_tensor_info(self.handle, i, buf, 128, C.byref(off), C.byref(n)), 'dsr_gen_tensor_info')
            out.append((buf.value.decode(), off.value, n.value))
        return out

    def __del__(self):
        try:
            if self.handle:
                lib.dsr_gen_plan_destroy(self.handle)
        except Exception:
            pass


class Generator(nn.Module):
    """``Generator(factor=8, residual_blocks_count=16)`` of generator.py:44-66."""

    max_chunk = 32        # images per library call (bounds the workspace: ~105 MB per 96 x 96 image at x8)
    max_plans = 4         # plans (one per batch / image size, each with its workspace) kept alive; least recently used go

    def __init__(self, factor: int = 8, residual_blocks_count: int = 16):
        super().__init__()
        if factor not in _SHUFFLES:
            # the reference leaves `pixel_shuffles` unbound for any other factor (generator.py:55-58) and fails
            raise NotImplementedError('dsr_b200.Generator: factor must be 8 or 16 (as in the reference)')
        self.factor, self.blocks = factor, residual_blocks_count
        self.conv1 = nn.Conv2d(3, 64, 9, 1, 4)
        self.prelu1 = nn.PReLU()
        self.residual_blocks = nn.Sequential(*[ResidualBlock() for _ in range(residual_blocks_count)])
        self.conv2 = nn.Conv2d(64, 64, 3, 1, 1)
        self.bn1 = nn.BatchNorm2d(64)
        self.pixel_shuffle_blocks = nn.Sequential(*[PixelShuffleBlock(64) for _ in range(_SHUFFLES[factor])])
        self.conv3 = nn.Conv2d(64, 3, 9, 1, 4)
        self.out = nn.Tanh()
        self._plans: Dict[Tuple[int, int, int, int], _GenPlan] = {}
        from .gan_train import FlatParams, NET_G
        self._fp = FlatParams(self, NET_G)

    # ---- training mode (train_GAN.py) ----------------------------------------------------------
    def _train_state(self, lr: torch.Tensor):
        from .gan_train import get_trainer
        B, _, h, w = lr.shape
        tr = get_trainer(B, h * self.factor, w * self.factor, lr.device, self.factor, self.blocks)
        self._fp.ensure(tr, lr.device)
        return tr, self._fp

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._fp.invalidate()
        for p in self.parameters():
            p.grad = None

    # ------------------------------------------------------------------------------------------
    def _state_key(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def _plan_for(self, batch: int, h: int, w: int, device: torch.device) -> _GenPlan:
        key = (batch, h, w, device.index if device.index is not None else torch.cuda.current_device())
        plan = self._plans.pop(key, None)
        if plan is None:
            while len(self._plans) >= self.max_plans:          # eval_GAN.py feeds images of many different sizes
                self._plans.pop(next(iter(self._plans)))
            plan = _GenPlan(self.factor, self.blocks, batch, h, w, device)
        self._plans[key] = plan                                # most recently used last
        skey = self._state_key()
        if plan.loaded_key != skey:
            sd = self.state_dict()
            parts = []
            for name, _off, numel in plan.names():
                t = sd[name]
                if t.numel() != numel:
                    raise RuntimeError(f'dsr_b200.Generator: {name} has {t.numel()} elements, expected {numel}')
                parts.append(t.detach().to(device=device, dtype=torch.float32).reshape(-1))
            flat = torch.cat(parts).contiguous()
            assert flat.numel() == lib.dsr_gen_state_numel(plan.handle)
            check(lib.dsr_gen_load_weights(plan.handle, flat.data_ptr(), _lib.stream_ptr()), 'dsr_gen_load_weights')
            plan.loaded_key = skey
        return plan

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError('dsr_b200.Generator needs a CUDA tensor (sm_100a); there is no CPU fallback')
        if self.training:                  # train_GAN.py: batch-statistics BatchNorm + backward (dsr_gant_*)
            if x.dim() != 4 or x.shape[1] != 3:
                raise ValueError('expected [B, 3, h, w]')
            from .gan_train import generator_train_forward
            return generator_train_forward(self, x)
        if not x.is_cuda:
            raise RuntimeError('dsr_b200.Generator needs a CUDA tensor (sm_100a); there is no CPU fallback')
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError('expected [B, 3, h, w]')
        B, _, h, w = x.shape
        f = self.factor
        xin = x.detach().to(torch.float32).contiguous()
        y = torch.empty((B, 3, f * h, f * w), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            for b0 in range(0, B, self.max_chunk):
                nb = min(self.max_chunk, B - b0)
                plan = self._plan_for(nb, h, w, x.device)
                check(lib.dsr_gen_forward(plan.handle, xin[b0:b0 + nb].data_ptr(), y[b0:b0 + nb].data_ptr(),
                                          _lib.stream_ptr()), 'dsr_gen_forward')
        return y
