"""Fixed Lanczos anti-aliasing downsampler: `Downsampler`, `get_kernel`.

Mirrors utils/downsampler.py:5-71 (Downsampler) and :73-134 (get_kernel) of the reference for the
configuration DIP.py:29 uses: kernel_type 'lanczos2' (also 'lanczos3'), phase 0.5,
preserve_size=True.  The reference evaluates a dense n_planes x n_planes strided convolution whose
weight is the 2-D table on the diagonal; the table is an outer product, so the CUDA kernels
(csrc/dsr_downsampler.cu) evaluate it separably per plane, clamp-indexed (ReplicationPad2d).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check


def get_kernel(factor, kernel_type, phase, kernel_width, support=None, sigma=None):
    """utils/downsampler.py:73-134, 'lanczos' branch with phase 0.5 (float64 numpy table)."""
    if kernel_type != 'lanczos' or phase != 0.5 or support not in (2, 3) or kernel_width != 2 * support * factor + 1:
        raise NotImplementedError("dsr_b200.get_kernel supports kernel_type='lanczos', phase=0.5, support 2 or 3, "
                                  'kernel_width = 2*support*factor + 1')
    k = 2 * support * factor
    buf = (C.c_double * (k * k))()
    rc = lib.dsr_lanczos_kernel(factor, support, buf, k * k)
    if rc != k:
        check(rc if rc < 0 else -1, 'dsr_lanczos_kernel')
    return np.frombuffer(buf, dtype=np.float64).reshape(k, k).copy()


class _Tables:
    def __init__(self, factor: int, support: int, H: int, W: int, device: torch.device):
        nbytes = lib.dsr_downsampler_table_bytes(factor, support, H, W)
        if nbytes == 0:
            raise ValueError(f'unsupported downsampler geometry factor={factor} H={H} W={W}')
        self.table = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.handle = C.c_void_p()
        check(lib.dsr_downsampler_create(C.byref(self.handle), factor, support, H, W, self.table.data_ptr(), nbytes,
                                         _lib.stream_ptr()), 'dsr_downsampler_create')

    def __del__(self):
        try:
            if self.handle:
                lib.dsr_downsampler_destroy(self.handle)
        except Exception:
            pass


class _DownsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mod):
        n, c, H, W = x.shape
        t = mod._tables_for(H, W, x.device)
        xc = x.detach().contiguous()
        oh, ow = mod.out_size(H, W)
        y = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
        check(lib.dsr_downsample_fwd(t.handle, xc.data_ptr(), y.data_ptr(), n * c, _lib.stream_ptr()),
              'dsr_downsample_fwd')
        ctx.t, ctx.shape = t, (n, c, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, c, H, W = ctx.shape
        g = gy.contiguous()
        gx = torch.empty((n, c, H, W), dtype=torch.float32, device=gy.device)
        check(lib.dsr_downsample_bwd(ctx.t.handle, g.data_ptr(), gx.data_ptr(), n * c, _lib.stream_ptr()),
              'dsr_downsample_bwd')
        return gx, None


class Downsampler(nn.Module):
    """Same constructor signature as utils/downsampler.py:9."""

    def __init__(self, n_planes, factor, kernel_type, phase=0, kernel_width=None, support=None, sigma=None,
                 preserve_size=False):
        super().__init__()
        assert phase in [0, 0.5], 'phase should be 0 or 0.5'
        if kernel_type not in ('lanczos2', 'lanczos3') or phase != 0.5 or not preserve_size:
            raise NotImplementedError("dsr_b200.Downsampler supports kernel_type 'lanczos2' / 'lanczos3' with "
                                      'phase=0.5 and preserve_size=True (the DIP.py:29 configuration)')
        self.n_planes, self.factor = int(n_planes), int(factor)
        self.support = 2 if kernel_type == 'lanczos2' else 3
        self.kernel = get_kernel(self.factor, 'lanczos', 0.5, 2 * self.support * self.factor + 1, support=self.support)
        # The reference builds an nn.Conv2d here and overwrites its weights (utils/downsampler.py:44-46); its random
        # initialisation consumes the global CPU generator between get_net and get_noise (DIP.py:29,32).  Draw the same
        # numbers so that a run with the same torch.manual_seed sees the same net_input / noise as the reference.
        nn.Conv2d(self.n_planes, self.n_planes, kernel_size=self.kernel.shape, stride=self.factor, padding=0)
        self.preserve_size = preserve_size
        self._tables: Dict[Tuple[int, int, str], _Tables] = {}

    def out_size(self, H: int, W: int) -> Tuple[int, int]:
        k = self.kernel.shape[0]
        pad = (k - self.factor) // 2
        return (H + 2 * pad - k) // self.factor + 1, (W + 2 * pad - k) // self.factor + 1

    def _tables_for(self, H: int, W: int, device: torch.device) -> _Tables:
        key = (H, W, str(device))
        t = self._tables.get(key)
        if t is None:
            t = _Tables(self.factor, self.support, H, W, device)
            self._tables[key] = t
        return t

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if not input.is_cuda:
            raise RuntimeError('dsr_b200.Downsampler runs on a CUDA device only (no CPU fallback)')
        if input.dim() != 4 or input.dtype != torch.float32 or input.shape[1] != self.n_planes:
            raise ValueError(f'expected float32 [N, {self.n_planes}, H, W], got {tuple(input.shape)} {input.dtype}')
        return _DownsampleFn.apply(input, self)
