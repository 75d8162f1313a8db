"""The skip network of the DIP step: `get_net` / `SkipNet`.

Mirrors models/DIP/__init__.py:8-18 (get_net) and models/DIP/skip.py:3-96 (skip) of the reference
as instantiated at DIP.py:169-174.  The returned object is an ``nn.Module`` whose parameter /
buffer tree has the reference's ``state_dict`` keys and shapes (children numbered from 1,
models/DIP/utils.py:5-8) and whose same-seed initialisation equals the reference's bit for bit
(the convolutions are drawn in the construction order of skip.py:41-92).  Its forward and
backward are single calls into libdsr_b200.so; parameters live in one flat device buffer that the
fused Adam step (dsr_b200.optim.optimize) updates in place.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check


class _Box(nn.Module):
    """Parameter-less container; children are registered under the reference's 1-based names."""


class _ConvLeaf(nn.Module):
    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        super().__init__()
        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias)


class _BnLeaf(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.zeros((), dtype=torch.long))


def _conv_box(weight, bias, child: str = '1') -> _Box:
    # conv() of models/DIP/utils.py:83-105 is Sequential(padder, Conv2d): the Conv2d is child '1'
    # (pad='zero': no padder module, the Conv2d is child '0')
    b = _Box()
    b.add_module(child, _ConvLeaf(weight, bias))
    return b


class _PlanState:
    """Per-resolution execution plan + workspace."""

    def __init__(self, H: int, W: int, input_depth: int, num_scales: int, n_out: int, device: torch.device,
                 flags: int = 0):
        self.handle = C.c_void_p()
        check(lib.dsr_plan_create_ex(C.byref(self.handle), H, W, input_depth, num_scales, n_out, flags),
              'dsr_plan_create_ex')
        nbytes = lib.dsr_plan_workspace_bytes(self.handle)
        self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        self.ws_ptr = base
        check(lib.dsr_plan_bind(self.handle, base, nbytes, _lib.stream_ptr()), 'dsr_plan_bind')
        self.H, self.W = H, W
        self.forward_id = 0

    def __del__(self):
        try:
            if self.handle:
                lib.dsr_plan_destroy(self.handle)
        except Exception:
            pass


class _SkipNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, anchor, net):
        plan = net._plan_for(z)
        out = torch.empty((1, net.n_channels, z.shape[2], z.shape[3]), dtype=torch.float32, device=z.device)
        zc = z.detach().contiguous()
        bn = net._bnflat.data_ptr() if net.training else None
        check(lib.dsr_net_forward(plan.handle, net._flat.data_ptr(), zc.data_ptr(), out.data_ptr(), bn,
                                  _lib.stream_ptr()), 'dsr_net_forward')
        if net.training:
            net._nbt += 1
        plan.forward_id += 1
        ctx.net, ctx.plan, ctx.forward_id = net, plan, plan.forward_id
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, gout):
        net, plan = ctx.net, ctx.plan
        if ctx.forward_id != plan.forward_id:
            raise RuntimeError('dsr_b200: backward() must follow the forward it belongs to (the plan keeps the '
                               'activations of the most recent forward only)')
        (out,) = ctx.saved_tensors
        g = gout.contiguous()
        fresh = not net._grads_valid or net._params[0].grad is None
        target = net._gflat if fresh else torch.empty_like(net._gflat)
        check(lib.dsr_net_backward(plan.handle, net._flat.data_ptr(), out.data_ptr(), g.data_ptr(), target.data_ptr(),
                                   _lib.stream_ptr()), 'dsr_net_backward')
        if not fresh:
            net._gflat += target                     # gradient accumulation across backward() calls
        if net._params[0].grad is not net._grad_views[0]:
            for p, gv in zip(net._params, net._grad_views):
                p.grad = gv
        net._grads_valid = True
        return None, None, None


class SkipNet(nn.Module):
    """Hourglass encoder/decoder with 4-channel skip branches (models/DIP/skip.py)."""

    def __init__(self, input_depth: int = 32, n_channels: int = 3, num_scales: int = 5, pad: str = 'reflection',
                 upsample_mode: str = 'bilinear'):
        super().__init__()
        self.input_depth, self.n_channels, self.num_scales = input_depth, n_channels, num_scales
        self.pad, self.upsample_mode = pad, upsample_mode
        # DSR_PLAN_PAD_ZERO = 1, DSR_PLAN_UP_NEAREST = 2 (include/dsr_b200.h)
        self._plan_flags = (1 if pad == 'zero' else 0) | (2 if upsample_mode == 'nearest' else 0)
        ci = '0' if pad == 'zero' else '1'          # index of the Conv2d inside conv()'s Sequential
        probe = C.c_void_p()
        check(lib.dsr_plan_create_ex(C.byref(probe), 2 << num_scales, 2 << num_scales, input_depth, num_scales,
                                     n_channels, self._plan_flags), 'dsr_plan_create_ex (configuration check)')
        self._layout = _read_layout(probe)
        lib.dsr_plan_destroy(probe)

        # --- initial values, drawn in the reference's construction order (skip.py:41-92) ---
        nd, ns = 128, 4
        init: Dict[str, torch.Tensor] = {}

        def draw(name, cin, cout, k):
            m = nn.Conv2d(cin, cout, k)
            init[name + '.weight'] = m.weight.detach()
            init[name + '.bias'] = m.bias.detach()

        cin = input_depth
        for i in range(num_scales):
            P = '1.1.7.' * i
            draw(P + '1.0.1.' + ci, cin, ns, 1)
            draw(P + '1.1.1.' + ci, cin, nd, 3)
            draw(P + '1.1.4.' + ci, nd, nd, 3)
            draw(P + '3.' + ci, ns + nd, nd, 3)
            draw(P + '6.' + ci, nd, nd, 1)
            cin = nd
        draw('9.' + ci, nd, n_channels, 1)

        # --- module tree with the reference's names ---
        def level(i: int) -> _Box:
            P = '1.1.7.' * i
            model, concat, skipb, deeper = _Box(), _Box(), _Box(), _Box()
            model.add_module('1', concat)
            concat.add_module('0', skipb)
            concat.add_module('1', deeper)
            skipb.add_module('1', _conv_box(init[P + f'1.0.1.{ci}.weight'], init[P + f'1.0.1.{ci}.bias'], ci))
            skipb.add_module('2', _BnLeaf(ns))
            deeper.add_module('1', _conv_box(init[P + f'1.1.1.{ci}.weight'], init[P + f'1.1.1.{ci}.bias'], ci))
            deeper.add_module('2', _BnLeaf(nd))
            deeper.add_module('4', _conv_box(init[P + f'1.1.4.{ci}.weight'], init[P + f'1.1.4.{ci}.bias'], ci))
            deeper.add_module('5', _BnLeaf(nd))
            if i + 1 < num_scales:
                deeper.add_module('7', level(i + 1))
            model.add_module('2', _BnLeaf(ns + nd))
            model.add_module('3', _conv_box(init[P + f'3.{ci}.weight'], init[P + f'3.{ci}.bias'], ci))
            model.add_module('4', _BnLeaf(nd))
            model.add_module('6', _conv_box(init[P + f'6.{ci}.weight'], init[P + f'6.{ci}.bias'], ci))
            model.add_module('7', _BnLeaf(nd))
            return model

        top = level(0)
        for name, child in top.named_children():
            self.add_module(name, child)
        self.add_module('9', _conv_box(init[f'9.{ci}.weight'], init[f'9.{ci}.bias'], ci))

        names = [n for n, _ in self.named_parameters()]
        if names != [n for n, _, _ in self._layout['params']]:
            raise RuntimeError('dsr_b200: parameter order of the module tree and of the plan disagree')
        self._params: List[nn.Parameter] = [p for _, p in self.named_parameters()]
        import weakref
        me = weakref.ref(self)
        for p in self._params:
            p._dsr_owner = me
        self._flat: Optional[torch.Tensor] = None
        self._gflat: Optional[torch.Tensor] = None
        self._bnflat: Optional[torch.Tensor] = None
        self._nbt: Optional[torch.Tensor] = None
        self._grad_views: List[torch.Tensor] = []
        self._grads_valid = False
        self._plans: Dict[Tuple[int, int], _PlanState] = {}

    # ------------------------------------------------------------------------------------------
    def _bn_leaves(self) -> List[_BnLeaf]:
        return [m for m in self.modules() if isinstance(m, _BnLeaf)]

    def _flatten(self, device: torch.device) -> None:
        """(Re)packs parameters and BatchNorm buffers into the flat device buffers the library
        works on and re-points the nn.Parameters / buffers at views of them."""
        lay = self._layout
        flat = torch.empty(lay['nparam'], dtype=torch.float32, device=device)
        gflat = torch.zeros(lay['nparam'], dtype=torch.float32, device=device)
        views = []
        with torch.no_grad():
            for p, (_, off, shape) in zip(self._params, lay['params']):
                n = p.numel()
                v = flat[off:off + n].view(shape)
                v.copy_(p.data)
                p.data = v
                p.grad = None
                views.append(gflat[off:off + n].view(shape))
            leaves = self._bn_leaves()
            bnflat = torch.empty(lay['nbn'], dtype=torch.float32, device=device)
            nbt = torch.empty(len(leaves), dtype=torch.long, device=device)
            for j, (leaf, (_, off, c)) in enumerate(zip(leaves, lay['bns'])):
                rm, rv = bnflat[off:off + c], bnflat[off + c:off + 2 * c]
                rm.copy_(leaf.running_mean)
                rv.copy_(leaf.running_var)
                nbt[j] = leaf.num_batches_tracked.to(device)
                leaf.running_mean, leaf.running_var, leaf.num_batches_tracked = rm, rv, nbt[j]
        self._flat, self._gflat, self._bnflat, self._nbt = flat, gflat, bnflat, nbt
        self._grad_views = views
        self._grads_valid = False

    def _is_flat(self, device: torch.device) -> bool:
        if self._flat is None or self._flat.device != device:
            return False
        base = self._flat.data_ptr()
        first, last = self._params[0], self._params[-1]
        off_last = self._layout['params'][-1][1]
        return first.data_ptr() == base and last.data_ptr() == base + 4 * off_last and first.device == device

    def _plan_for(self, z: torch.Tensor) -> _PlanState:
        key = (int(z.shape[2]), int(z.shape[3]))
        plan = self._plans.get(key)
        if plan is None or plan.workspace.device != z.device:
            plan = _PlanState(key[0], key[1], self.input_depth, self.num_scales, self.n_channels, z.device,
                              self._plan_flags)
            self._plans[key] = plan
        return plan

    def flat_buffers(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(parameters, gradients) as flat fp32 device tensors -- what the fused Adam works on."""
        return self._flat, self._gflat

    def zero_grad(self, set_to_none: bool = True) -> None:  # noqa: D401
        self._grads_valid = False
        if set_to_none:
            for p in self._params:
                p.grad = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError('dsr_b200.SkipNet runs on a CUDA device only (no CPU fallback): call .to("cuda") on '
                               'the network and its input')
        if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != self.input_depth or x.dtype != torch.float32:
            raise ValueError(f'expected a float32 input of shape [1, {self.input_depth}, H, W], got {tuple(x.shape)} '
                             f'{x.dtype}')
        if x.requires_grad:
            raise NotImplementedError("optimising over the input (opt_over='input') is not supported")
        if not self.training:
            raise NotImplementedError('eval-mode BatchNorm is not supported: the DIP loop never leaves train mode '
                                      '(DIP.py:60,102)')
        if not self._is_flat(x.device):
            self._flatten(x.device)
        return _SkipNetFn.apply(x, self._params[0], self)

    # --- introspection used by the tests ---
    def debug_tensor(self, name: str, hw: Tuple[int, int]) -> torch.Tensor:
        plan = self._plans[hw]
        ptr, kind, padded, H, W, Cc = C.c_void_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib.dsr_plan_tensor(plan.handle, name.encode(), C.byref(ptr), C.byref(kind), C.byref(padded), C.byref(H),
                                  C.byref(W), C.byref(Cc)), f'dsr_plan_tensor({name})')
        # kind 3: 64-bit fixed-point accumulators (csrc/dsr_acc.cuh, forward scale 2^-20), returned as float64
        dt = {0: torch.float16, 1: torch.bfloat16, 2: torch.float32, 3: torch.int64}[kind.value]
        hh, ww = (H.value + 2, W.value + 2) if padded.value else (H.value, W.value)
        t = torch.empty((hh, ww, Cc.value), dtype=dt, device=self._flat.device)
        check(lib.dsr_debug_copy(t.data_ptr(), ptr, t.numel() * t.element_size(), _lib.stream_ptr()), 'dsr_debug_copy')
        return t.double() / float(1 << 20) if kind.value == 3 else t

    def set_debug_conv(self, on: bool) -> None:
        for plan in self._plans.values():
            check(lib.dsr_plan_set_debug_conv(plan.handle, int(on)))

    def last_launches(self, hw: Tuple[int, int]) -> int:
        return lib.dsr_plan_last_launches(self._plans[hw].handle)


def _read_layout(handle) -> dict:
    n = lib.dsr_plan_num_params(handle)
    name = C.create_string_buffer(128)
    off, nd, shape = C.c_longlong(), C.c_int(), (C.c_int * 4)()
    params = []
    for i in range(n):
        check(lib.dsr_plan_param_info(handle, i, name, 128, C.byref(off), C.byref(nd), shape))
        params.append((name.value.decode(), off.value, tuple(shape[j] for j in range(nd.value))))
    bns = []
    ch = C.c_int()
    for i in range(lib.dsr_plan_num_bn(handle)):
        check(lib.dsr_plan_bn_info(handle, i, name, 128, C.byref(off), C.byref(ch)))
        bns.append((name.value.decode(), off.value, ch.value))
    return dict(params=params, bns=bns, nparam=lib.dsr_plan_param_numel(handle), nbn=lib.dsr_plan_bn_numel(handle))


def get_net(input_depth, NET_TYPE, pad, upsample_mode, n_channels=3, act_fun='LeakyReLU', skip_n33d=128,
            skip_n33u=128, skip_n11=4, num_scales=5, downsample_mode='stride'):
    """Same signature as models/DIP/__init__.py:8.  Supported: the configuration DIP.py:169-174
    builds (NET_TYPE 'skip', pad 'reflection', upsample 'bilinear', LeakyReLU, 128/128/4 channels,
    stride downsampling) and the two other padding / upsampling options of the signature, pad='zero'
    (models/DIP/utils.py:96-102) and upsample_mode='nearest' (models/DIP/skip.py:77); anything else raises
    NotImplementedError -- there is no fallback."""
    def _all(v, want):
        return (v == want) if isinstance(v, int) else all(x == want for x in v) and len(v) == num_scales
    ok = (NET_TYPE == 'skip' and pad in ('reflection', 'zero') and upsample_mode in ('bilinear', 'nearest')
          and act_fun == 'LeakyReLU'
          and downsample_mode == 'stride' and _all(skip_n33d, 128) and _all(skip_n33u, 128) and _all(skip_n11, 4))
    if not ok:
        raise NotImplementedError(
            'dsr_b200.get_net supports NET_TYPE="skip", pad="reflection"|"zero", '
            'upsample_mode="bilinear"|"nearest", act_fun="LeakyReLU", skip_n33d=skip_n33u=128, skip_n11=4, '
            'downsample_mode="stride"')
    return SkipNet(input_depth, n_channels, num_scales, pad=pad, upsample_mode=upsample_mode)
