"""Optimisation helpers of the DIP loop: `optimize`, `get_params`, `get_noise`, `fill_noise`.

Mirrors utils/DIP.py:7-105 of the reference for the branches DIP.py reaches: optimizer 'adam'
(:33-40), opt_over 'net' (:55-58), method 'noise' with noise_type 'u' / 'n' (:70-77, :92-96).
`optimize` owns the optimiser: one fused multi-tensor Adam pass (torch.optim.Adam defaults,
csrc/dsr_elem.cu adam_kernel) over the network's flat parameter buffer.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import lib, check
from .net import SkipNet


def _owner(parameters) -> SkipNet:
    params = list(parameters)
    ref = getattr(params[0], '_dsr_owner', None) if params else None
    net = ref() if ref is not None else None
    if net is None or len(net._params) != len(params) or any(a is not b for a, b in zip(net._params, params)):
        raise NotImplementedError('dsr_b200.optimize expects exactly the parameter list of one dsr_b200 SkipNet '
                                  "(get_params('net', net, net_input)); other parameter sets are not supported")
    return net


def optimize(optimizer_type, parameters, closure, learning_rate, num_iter):
    """utils/DIP.py:7-42, 'adam' branch: num_iter x (clear grads, closure(), Adam step)."""
    if optimizer_type != 'adam':
        raise NotImplementedError("dsr_b200.optimize supports optimizer_type='adam' only (DIP.py:99)")
    net = _owner(parameters)
    m = v = None
    beta1, beta2, eps = 0.9, 0.999, 1e-8          # torch.optim.Adam defaults (utils/DIP.py:34)
    for t in range(1, int(num_iter) + 1):
        net.zero_grad(set_to_none=False)           # optimizer.zero_grad(): stale gradients are never accumulated
        closure()
        flat, gflat = net.flat_buffers()
        if flat is None or not net._grads_valid:
            raise RuntimeError('dsr_b200.optimize: closure() did not run forward + backward of the network')
        if m is None or m.data_ptr() == 0 or m.device != flat.device or m.numel() != flat.numel():
            m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        check(lib.dsr_adam_step(flat.data_ptr(), gflat.data_ptr(), m.data_ptr(), v.data_ptr(), flat.numel(),
                                float(learning_rate), beta1, beta2, eps, t, _lib.stream_ptr()), 'dsr_adam_step')
    net.zero_grad(set_to_none=True)                # utils/DIP.py:39


def get_params(opt_over, net, net_input, downsampler=None):
    """utils/DIP.py:44-68; only 'net' is supported."""
    params = []
    for opt in opt_over.split(','):
        if opt == 'net':
            params += [x for x in net.parameters()]
        else:
            raise NotImplementedError(f"dsr_b200.get_params: opt_over='{opt}' is not supported (only 'net')")
    return params


def fill_noise(x, noise_type):
    """utils/DIP.py:70-77."""
    if noise_type == 'u':
        x.uniform_()
    elif noise_type == 'n':
        x.normal_()
    else:
        assert False


def get_noise(input_depth, method, spatial_size, noise_type='u', var=1. / 10):
    """utils/DIP.py:79-105, method 'noise'.

    The values are drawn on the CPU generator exactly as the reference does (same seed -> same
    z), then moved to the current CUDA device so that the closure's `noise.normal_()` and add
    (DIP.py:52) run on the device and `.to(device)` (DIP.py:57) is free.  Set DSR_NOISE_DEVICE=cpu
    to keep the tensor on the host (parity runs that must consume the CPU normal stream).
    """
    if isinstance(spatial_size, int):
        spatial_size = (spatial_size, spatial_size)
    if method != 'noise':
        raise NotImplementedError("dsr_b200.get_noise supports method='noise' only (DIP.py:32)")
    net_input = torch.zeros([1, input_depth, spatial_size[0], spatial_size[1]])
    fill_noise(net_input, noise_type)
    net_input *= var
    if os.environ.get('DSR_NOISE_DEVICE', 'cuda') != 'cpu' and torch.cuda.is_available():
        net_input = net_input.to(torch.device('cuda', torch.cuda.current_device()))
    return net_input
