"""Per-image DIP super-resolution drivers.

`DIP_ISR` follows the reference's per-image driver (DIP.py:22-123) through the public call
surface -- Downsampler, get_noise, get_params, optimize and a per-iteration closure doing
forward, downsample, MSE, backward and the two blocking device-to-host reads of DIP.py:90-91.

`dip_sr_fused` is the same optimisation with every iteration enqueued by ONE library call
(dsr_dip_step: device Philox perturbation -> forward -> Lanczos downsample + MSE -> backward ->
Adam), replayed as a CUDA graph (dsr_dip_run), with no host synchronisation inside the loop; the per-iteration
losses stay on the device.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import lib, check, StepBuffers
from .downsampler import Downsampler
from .net import SkipNet
from .optim import get_noise, get_params, optimize


def DIP_ISR(net, LR_image, HR_image, scale_factor, training_config, train_log_freq, psnr, ssim, lpips, device):
    """Same signature and return value as DIP.py:22: (resolved image [1,3,H,W] on `device`,
    {'psnrs', 'ssims', 'lpipss'}).  `psnr` / `ssim` / `lpips` are metric callables as in the
    reference (torchmetrics objects there); any of them may be None to skip that metric."""
    import torch.nn.functional as F
    mse = torch.nn.MSELoss()                                                  # DIP.py:26
    downsampler = Downsampler(n_planes=3, factor=scale_factor, kernel_type='lanczos2', phase=0.5,
                              preserve_size=True).to(device)                   # DIP.py:29
    net_input = get_noise(32, 'noise', (HR_image.shape[1], HR_image.shape[2])).detach()   # DIP.py:32
    net_input_saved = net_input.detach().clone()
    noise = net_input.detach().clone()
    LR_image = LR_image.unsqueeze(0).to(device).detach()
    HR_image = HR_image.unsqueeze(0).to(device).detach()
    state = {'iter': 0, 'net_input': net_input}
    psnrs, ssims, lpipss = [], [], []
    sigma = training_config['reg_noise_std']

    def closure():
        z = state['net_input']
        if sigma > 0:
            z = net_input_saved + (noise.normal_() * sigma)                   # DIP.py:52
        start_time = time.time()
        z = z.to(device)                                                      # DIP.py:57
        state['net_input'] = z
        out_HR = net(z)                                                       # DIP.py:60
        out_LR = downsampler(out_HR)                                          # DIP.py:62
        total_loss = mse(out_LR, LR_image)                                    # DIP.py:65
        total_loss.backward()                                                 # DIP.py:68
        if state['iter'] % train_log_freq == 0 and (psnr or ssim or lpips):   # DIP.py:71-87
            if psnr is not None:
                psnrs.append(psnr(out_HR, HR_image).item())
            if ssim is not None:
                ssims.append(ssim(out_HR, HR_image).item())
            if lpips is not None:
                lpipss.append(lpips(F.normalize(out_HR, dim=0), F.normalize(HR_image, dim=0)).item())
            print(f"Iteration {state['iter'] + 1}/{training_config['num_iter']}: "
                  f"PSNR {psnrs[-1] if psnrs else None} SSIM {ssims[-1] if ssims else None} "
                  f"LPIPS {lpipss[-1] if lpipss else None} ({time.time() - start_time:.3f} s)")
        state['iter'] += 1
        out_HR.detach().cpu()                                                 # DIP.py:90-91 (blocking reads)
        out_LR.detach().cpu()
        return total_loss

    params = get_params('net', net, state['net_input'])                       # DIP.py:98
    optimize('adam', params, closure, training_config['learning_rate'], training_config['num_iter'])
    resolved_image = net(state['net_input'].to(device)).detach()              # DIP.py:102 (last perturbed input)
    downsampler.cpu()
    net.cpu()                                                                 # DIP.py:105-109
    return resolved_image, {'psnrs': psnrs, 'ssims': ssims, 'lpipss': lpipss}


def dip_sr_fused(net: SkipNet, LR_image: torch.Tensor, hr_size: Tuple[int, int], scale_factor: int,
                 training_config: Dict, device, seed: int = 0, net_input: Optional[torch.Tensor] = None,
                 keep_on_device: bool = True, callback=None, callback_from: int = 1):
    """One image, `num_iter` fused iterations.  Returns (resolved image [1,3,H,W], losses [num_iter]
    device tensor).  `training_config`: 'learning_rate', 'num_iter', 'reg_noise_std' as in DIP.py:316-324.
    `callback(t, out_hr)` (optional) is called after iterations t >= callback_from with the network output of that
    iteration (the tensor is overwritten by the next iteration), e.g. to log PSNR as DIP.py:71-87 does."""
    _lib.require_cuda()
    device = torch.device(device if not isinstance(device, int) else f'cuda:{device}')
    H, W = int(hr_size[0]), int(hr_size[1])
    num_iter = int(training_config['num_iter'])
    # a non-default stream: the legacy default stream cannot be captured into a CUDA graph
    run_stream = torch.cuda.Stream(device=device)
    run_stream.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.device(device), torch.cuda.stream(run_stream):
        net.to(device)
        z_saved = (net_input if net_input is not None else get_noise(net.input_depth, 'noise', (H, W))).detach()
        z_saved = z_saved.to(device).contiguous()
        z = z_saved.clone()
        if not net._is_flat(device):
            net._flatten(device)
        plan = net._plan_for(z)
        ds = Downsampler(n_planes=net.n_channels, factor=scale_factor, kernel_type='lanczos2', phase=0.5,
                         preserve_size=True)
        tables = ds._tables_for(H, W, device)
        oh, ow = ds.out_size(H, W)
        lr = LR_image.to(device).reshape(net.n_channels, oh, ow).contiguous().float()
        flat, gflat = net.flat_buffers()
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        f32 = dict(dtype=torch.float32, device=device)
        out_hr = torch.empty((1, net.n_channels, H, W), **f32)
        out_lr = torch.empty((net.n_channels, oh, ow), **f32)
        g_lr = torch.empty_like(out_lr)
        g_hr = torch.empty_like(out_hr)
        losses = torch.zeros(max(num_iter, 1), **f32)
        b = StepBuffers(flat.data_ptr(), gflat.data_ptr(), m.data_ptr(), v.data_ptr(), net._bnflat.data_ptr(),
                        z_saved.data_ptr(), z.data_ptr(), lr.data_ptr(), out_hr.data_ptr(), out_lr.data_ptr(),
                        g_lr.data_ptr(), g_hr.data_ptr(), losses.data_ptr())
        stream = _lib.stream_ptr()
        lr_rate, sigma = float(training_config['learning_rate']), float(training_config['reg_noise_std'])
        # the whole loop in one call: iteration 1 runs eagerly, the rest replay one captured CUDA graph
        n_bulk = num_iter if callback is None else max(0, min(num_iter, int(callback_from) - 1))
        check(lib.dsr_dip_run(plan.handle, tables.handle, C.byref(b), lr_rate, sigma, seed, 1, n_bulk, stream),
              'dsr_dip_run')
        for t in range(n_bulk + 1, num_iter + 1):
            check(lib.dsr_dip_run(plan.handle, tables.handle, C.byref(b), lr_rate, sigma, seed, t, 1, stream),
                  'dsr_dip_run')
            callback(t, out_hr)
        net._nbt += num_iter
        # final resolved image: net(last perturbed input), BatchNorm still in train mode (DIP.py:102)
        check(lib.dsr_net_forward(plan.handle, flat.data_ptr(), z.data_ptr(), out_hr.data_ptr(),
                                  net._bnflat.data_ptr(), stream), 'dsr_net_forward')
        net._nbt += 1
        plan.forward_id += 1
        net._fused_keepalive = (m, v, z, z_saved, lr, out_lr, g_lr, g_hr, tables, ds)
    caller = torch.cuda.current_stream(device)
    caller.wait_stream(run_stream)
    # the returned tensors were allocated on run_stream: tell the caching allocator that the caller's stream uses them
    # too, or their blocks could be handed to later run_stream allocations while the caller still reads them
    out_hr.record_stream(caller)
    losses.record_stream(caller)
    return (out_hr if keep_on_device else out_hr.cpu()), losses
