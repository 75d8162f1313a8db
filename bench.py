#!/usr/bin/env python
"""bench.py -- DIP iterations/s of the B200-native step (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--impl ours|reference]

A "step" is ONE Deep-Image-Prior iteration at 512x512, factor 4 (BASELINE.json configs[1]): input perturbation,
skip-net forward, Lanczos downsample, MSE, full backward, Adam -- the closure of DIP.py:47-95 plus
utils/DIP.py:35-38.

* value      device-resident: the K iterations are one dsr_dip_run() call = one CUDA-graph launch per iteration, noise
             drawn on the device, nothing crosses PCIe.  K steps timed with CUDA events between
             barrier + synchronize, max over ranks; N ranks each optimise their own image (weak scaling, no
             collective) and value = N*K / time.
* e2e        the same iteration through the reference-facing Python call surface (get_net, Downsampler,
             get_params, optimize and a closure written like DIP.py:47-95) with HOST buffers: the step's perturbed
             input z comes from pinned host memory (H2D inside the timed region, as DIP.py:57 does) and out_HR /
             out_LR / loss are read back every step (DIP.py:90-91).
* roofline   the tensor-core kernel class with the LARGEST time per iteration (today wgrad_halo_kernel): algorithmic
             FLOPs / time per launch, summed over profiled iterations (all 5 levels, the latency-bound 16x16..64x64
             launches included; in-kernel %globaltimer stamps, dependencies satisfied -> last CTA done), against the
             BURST bf16 peak of MEASURED_PEAKS.json.  Every class is listed under roofline.classes (the 1x1 launches,
             64 FLOP per byte, against the HBM roof), the largest launch of the two big classes and the whole step too.
* cpu_baseline / --impl reference   the UNMODIFIED reference modules (baseline/_ref, staged by __graft_entry__.build())
             with a closure written like DIP.py:47-69 and torch.optim.Adam on the box's host cores, all threads, same
             workload; gpu_eager_baseline = the same modules on stock PyTorch CUDA eager (fp32, TF32 off).
* secondary / secondary_gan_train   short single-GPU runs of the two other BASELINE workloads (generator inference,
             configs[3]; SRGAN training step, configs[4]); their full lines: --workload gan_eval / gan_train (the
             latter data-parallel under torchrun, NCCL all-reduce of the flat gradients).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'deep-super-resolution_b200')
REF = os.path.join(ROOT, 'baseline', '_ref')          # unmodified reference checkout, staged by __graft_entry__.build()
# The reference arm imports the reference's OWN models/ and utils/ packages; our arm imports the drop-in ones.  The two
# share module names, so the path order is decided by --impl before anything is imported.
_REFERENCE_ARM = '--impl' in sys.argv and sys.argv[sys.argv.index('--impl') + 1:][:1] == ['reference']
for p in ((REF, ROOT) if _REFERENCE_ARM and os.path.isdir(REF) else (PKG, ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

FLOPS_PER_ITER = {512: 460.07e9, 256: 115.02e9}      # convs, fwd + dgrad + wgrad, true K
FACTOR = 4
LR_RATE, SIGMA = 0.01, 0.05                          # DIP.py:318,323


def peaks():
    """Roofline denominators.  Kernel classes are timed per launch (CUDA events, the GPU otherwise idle), so the BURST
    bf16 figure is the tensor peak; the whole-step figure is also quoted against the sustained one."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tflops=d.get('bf16_tflops', d.get('bf16_tflops_sustained')),
                    tflops_sustained=d.get('bf16_tflops_sustained', d.get('bf16_tflops')),
                    hbm=d.get('hbm_gbs'), src='measured (MEASURED_PEAKS.json: burst bf16, HBM copy)')
    return dict(tflops=1590.0, tflops_sustained=1590.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


def ncu_traffic(name):
    """DRAM bytes (read + write) of one launch from a committed `ncu --set full` capture under profiles/ (ARCHIVED, not
    measured by this run -- ncu cannot run inside the bench); None if the file is missing."""
    import csv
    path = os.path.join(ROOT, 'profiles', name)
    try:
        rows = list(csv.reader(open(path)))
        hdr, units, row = rows[0], rows[1], rows[2]
        tot = 0.0
        for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(name)
            scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
            tot += float(row[i]) * scale
        return tot
    except Exception:
        return None


def synthetic_pair(index, size):
    """HR / LR of image `index` (SURVEY.md 8d recipe).  The LR image is produced with the CUDA Lanczos downsampler
    so that bench.py does not depend on the oracle for its own arm."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1000 + index)
    low = torch.rand(1, 3, size // 8, size // 8, generator=g)
    field = F.interpolate(low, size=(size, size), mode='bicubic', align_corners=False)[0]
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    checker = (((yy // 16) + (xx // 16)) % 2).float() - 0.5
    return (field * 0.7 + 0.15 + 0.3 * checker).clamp(0, 1)


class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), f'--query-gpu={self.QUERY}',
                                       '--format=csv,noheader,nounits', '-lms', '50'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def nsamples(self):
        try:
            self.f.flush()
            return sum(1 for r in open(self.f.name) if r.strip())
        except Exception:
            return 0

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(',') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            sm.sort()
            out = dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# -------------------------------------------------------------------------------------------------
# Reference arm: the UNMODIFIED reference modules (baseline/_ref: models/DIP, utils/downsampler.py, utils/DIP.py)
# driven by a closure that repeats DIP.py:47-69 -- on the host cores (the CPU baseline) or, harness only, on the GPU
# with stock PyTorch CUDA eager, fp32, TF32 off (the like-for-like GPU comparator of BASELINE.md section 3).
# Falls back to the oracle port (oracle/dip_oracle.py) only where baseline/_ref was not staged.
# -------------------------------------------------------------------------------------------------
WORKLOAD = 'DIP 4x SR of one synthetic {s}x{s} image, skip net fwd+bwd+Adam (BASELINE configs[1])'


def reference_steps(size, warmup, steps, budget_s, device='cpu', predrawn=False):
    """it/s (median step) of the reference's own modules; returns (it/s, steps timed, cores, kind)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if not os.path.isdir(REF) or not _REFERENCE_ARM:
        its, n, cores = port_steps(size, warmup, steps, budget_s)
        return its, n, cores, 'port'
    from models.DIP import get_net                        # the reference's own files (REF is first on sys.path)
    from utils.downsampler import Downsampler
    from utils.DIP import get_noise, get_params, optimize
    import models.DIP as _m
    assert os.path.realpath(_m.__file__).startswith(os.path.realpath(REF)), _m.__file__
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device(device)
    hr = synthetic_pair(0, size)
    ds = Downsampler(n_planes=3, factor=FACTOR, kernel_type='lanczos2', phase=0.5, preserve_size=True)
    with torch.no_grad():
        lr_img = ds(hr.unsqueeze(0)).to(dev)
    ds = ds.to(dev)
    torch.manual_seed(0)
    net = get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                  upsample_mode='bilinear').to(dev)                                   # DIP.py:169-174
    net_input = get_noise(32, 'noise', (size, size)).detach()                        # DIP.py:32 (a CPU tensor)
    saved, noise = net_input.detach().clone(), net_input.detach().clone()
    mse = torch.nn.MSELoss()
    times, t_begin = [], time.time()
    state = {'t0': 0.0}

    ring = [(saved + torch.randn(saved.shape) * SIGMA).pin_memory() for _ in range(4)] if predrawn else None

    def closure():
        if predrawn:      # comparator for `e2e`: the perturbed inputs are drawn ahead, like the ring our e2e arm reads
            z = ring[len(times) % 4]
        else:
            z = saved + (noise.normal_() * SIGMA)                                     # DIP.py:52 (host RNG)
        z = z.to(dev)                                                                 # DIP.py:57
        out_hr = net(z)
        out_lr = ds(out_hr)
        loss = mse(out_lr, lr_img)
        loss.backward()
        out_hr.detach().cpu()                                                         # DIP.py:90-91
        out_lr.detach().cpu()
        return loss

    params = get_params('net', net, net_input)
    optim = torch.optim.Adam(params, lr=LR_RATE)                                      # utils/DIP.py:34
    for i in range(warmup + steps):
        t0 = time.time()
        optim.zero_grad()                                                             # utils/DIP.py:36-38
        closure()
        optim.step()
        if dev.type == 'cuda':
            torch.cuda.synchronize()
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_begin > budget_s and len(times) >= 2:
            break
    times.sort()
    return 1.0 / times[len(times) // 2], len(times), cores, 'reference'


def port_steps(size, warmup, steps, budget_s):
    """Times DIP iterations of the CPU restatement (fp32, all host threads).  Returns (it/s, steps timed, cores)."""
    from oracle import dip_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = O.init_params()
    lr_img, hr = O.synthetic_pair(0, size)
    z_saved = torch.zeros(1, 32, size, size).uniform_() * 0.1
    noise = z_saved.clone()
    keys = O.param_keys(sd)
    adam = O.AdamState(keys, sd, LR_RATE)
    times = []
    t_begin = time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        z = z_saved + noise.normal_() * SIGMA                                  # DIP.py:52 (host RNG, as the reference)
        loss, out, grads = O.step_loss_and_grads(sd, z, lr_img.unsqueeze(0), FACTOR)
        adam.step(sd, grads)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_begin > budget_s and len(times) >= 2:
            break
    times.sort()
    med = times[len(times) // 2]
    return 1.0 / med, len(times), cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    dev = args.ref_device
    its, n, cores, kind = reference_steps(args.size, args.warmup, args.steps, budget_s=150.0, device=dev,
                                          predrawn=args.ref_predrawn)
    what = ('unmodified reference modules (baseline/_ref: models/DIP, utils/downsampler.py, utils/DIP.py) + a '
            'DIP.py:47-69 closure + torch.optim.Adam' if kind == 'reference' else 'oracle/dip_oracle.py on torch CPU')
    where = f'{cores} host cores' if dev == 'cpu' else 'stock PyTorch CUDA eager, fp32, TF32 off (harness only)'
    line = {
        'impl': 'reference', 'metric': 'DIP iters/s (512^2 4x SR)', 'value': its, 'unit': 'it/s', 'n_gpus': args.gpus,
        'steps': n, 'warmup': args.warmup, 'ms_per_step': 1000.0 / its, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD.format(s=args.size), 'size': args.size, 'factor': FACTOR},
        'cpu_baseline': {'value': its, 'unit': 'it/s', 'cores': cores, 'kind': kind, 'device': dev,
                         'sample': f'{n} full iterations at {args.size}^2 (median step time), {what}, on {where}'},
        'e2e': {'value': its, 'unit': 'it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def reference_subprocess(size, device, steps, warmup, predrawn=False):
    """Runs `bench.py --impl reference` in a child process (the reference's models/ and utils/ packages share their
    names with the drop-in ones, so the two cannot be imported side by side) and returns its JSON line."""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--ref-device', device, '--size', str(size),
           '--steps', str(steps), '--warmup', str(warmup)] + (['--ref-predrawn'] if predrawn else [])
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE')}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env).stdout
        return json.loads(out.strip().splitlines()[-1])
    except Exception as e:            # noqa: BLE001
        return {'unavailable': repr(e)[:200]}


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import ctypes as C
    import torch.distributed as dist
    import dsr_b200
    from dsr_b200._lib import lib, check, StepBuffers
    from dsr_b200 import _lib

    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    # everything runs on one non-default stream (the legacy default stream cannot be captured into a CUDA graph)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=int(os.environ.get('DSR_BENCH_PRIO', '0'))))
    size = args.size
    H = W = size

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: image `rank`, fresh same-seed network (DIP.py:169) ----
    hr = synthetic_pair(rank, size)
    ds = dsr_b200.Downsampler(3, FACTOR, 'lanczos2', phase=0.5, preserve_size=True)
    lr_img = ds(hr.unsqueeze(0).to(dev))[0].contiguous()
    torch.manual_seed(rank)
    net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                           upsample_mode='bilinear').to(dev)
    z_saved = dsr_b200.get_noise(32, 'noise', (H, W)).to(dev).contiguous()
    z = z_saved.clone()
    net(z)                                   # flattens parameters, builds the plan + workspace
    net.zero_grad()
    plan = net._plans[(H, W)]
    tables = ds._tables_for(H, W, dev)
    oh, ow = ds.out_size(H, W)
    flat, gflat = net.flat_buffers()
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    f32 = dict(dtype=torch.float32, device=dev)
    out_hr = torch.empty((1, 3, H, W), **f32)
    out_lr, g_lr, g_hr = torch.empty((3, oh, ow), **f32), torch.empty((3, oh, ow), **f32), torch.empty((1, 3, H, W), **f32)
    total = args.warmup + args.steps + 64 + 4096      # + the untimed clock-sampling loop (<= 200 x 20 iterations)
    losses = torch.zeros(total, **f32)
    b = StepBuffers(flat.data_ptr(), gflat.data_ptr(), m.data_ptr(), v.data_ptr(), net._bnflat.data_ptr(),
                    z_saved.data_ptr(), z.data_ptr(), lr_img.data_ptr(), out_hr.data_ptr(), out_lr.data_ptr(),
                    g_lr.data_ptr(), g_hr.data_ptr(), losses.data_ptr())
    stream = _lib.stream_ptr()
    t_iter = [0]

    def step(n=1):
        """n iterations through dsr_dip_run (one CUDA-graph launch per iteration after the first call)"""
        check(lib.dsr_dip_run(plan.handle, tables.handle, C.byref(b), LR_RATE, SIGMA, 1234 + rank, t_iter[0] + 1, n,
                              stream), 'dsr_dip_run')
        t_iter[0] += n

    # ---- device-resident throughput ----
    step(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = None
    if sampler:
        # nvidia-smi needs a few hundred ms to produce its first line and the timed region is ~35 ms: keep the SAME
        # load running (untimed) until a handful of samples exist, so that the clocks line describes the GPU under
        # this workload (the window is stated in the line)
        t_end, loops = time.time() + 3.0, 0
        while sampler.nsamples() < 5 and time.time() < t_end and loops < 200:
            loops += 1
            step(20)
            torch.cuda.synchronize()
        clocks = sampler.stop()
        clocks['window'] = 'timed region + the same iteration loop kept running (untimed) until >= 5 samples at 50 ms'
    launches_per_step = lib.dsr_plan_last_launches(plan.handle)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    loss_first, loss_last = float(losses[args.warmup]), float(losses[args.warmup + args.steps - 1])

    # ---- roofline: profiled passes (CUDA events around every tensor-core launch, the host enqueueing ahead of the
    #      GPU behind a busy-wait, so the intervals are GPU time).  Four kernel classes; the DOMINANT one is the class
    #      with the largest time per iteration.  Per-launch timings -> burst bf16 peak.
    roof = None
    if rank == 0:
        check(lib.dsr_plan_set_profile(plan.handle, 1))
        nprof = 3
        step(nprof)
        torch.cuda.synchronize()
        pk = peaks()
        # conv-side algorithmic bytes per iteration of each class at this size (every operand / result read or
        # written once at its storage dtype), used for the HBM-bound 1x1 class: 2 x 128 channels x 2 B per pixel and pass
        res = {}
        for cls, name, what in (
                (0, 'conv_halo2_kernel', '3x3 stride-1 fprop + dgrad, halo-tile implicit GEMM, tcgen05 cta_group::2'),
                (1, 'wgrad_halo_kernel', 'all weight gradients (split-K over pixels, both operands MN-major)'),
                (2, 'conv_gemm_kernel', 'stride-2 layers + 32-channel first layer, generic implicit GEMM'),
                (3, 'conv_halo2_kernel_1x1', '1x1 launches of the halo-tile kernel: 64 FLOP/B, HBM-bound')):
            msx, fl, n = C.c_double(), C.c_double(), C.c_int()
            check(lib.dsr_plan_profile_read(plan.handle, cls, C.byref(msx), C.byref(fl), C.byref(n)))
            tf = (fl.value / (msx.value * 1e-3) / 1e12) if msx.value else 0.0
            res[name] = dict(what=what, ms_per_step=msx.value / nprof, launches_per_step=n.value // nprof,
                             gflop_per_step=fl.value / nprof / 1e9, tflops=tf, bound='tensor',
                             frac=tf / pk['tflops'] if pk['tflops'] else None)
        one = res['conv_halo2_kernel_1x1']
        one_gbs = (one['gflop_per_step'] / 64.0) / (one['ms_per_step'] * 1e-3) if one['ms_per_step'] else 0.0
        one.update({'bound': 'hbm', 'achieved_gbs': one_gbs, 'peak_gbs': pk['hbm'],
                    'frac': one_gbs / pk['hbm'] if pk['hbm'] else None})
        tops = {}
        for cls, name in ((0, 'conv_halo2_kernel'), (1, 'wgrad_halo_kernel')):
            tms, tfl = C.c_double(), C.c_double()
            check(lib.dsr_plan_profile_top(plan.handle, cls, C.byref(tms), C.byref(tfl)))
            tf = (tfl.value / (tms.value * 1e-3) / 1e12) if tms.value else 0.0
            tops[name] = {'gflop': tfl.value / 1e9, 'us': tms.value * 1e3, 'tflops': tf,
                          'frac': tf / pk['tflops'] if pk['tflops'] else None}
        check(lib.dsr_plan_set_profile(plan.handle, 0))
        dom = max(res, key=lambda k: res[k]['ms_per_step'])
        d = res[dom]
        arch = {'wgrad_halo_kernel': ('r02_wgrad_full_raw.csv', 'r01_wgrad_full_raw.csv'),
                'conv_halo2_kernel': ('r02_halo2_full_raw.csv', 'r01_halo2_full_raw_v3.csv')}.get(dom, ())
        traffic, tsrc = None, None
        for f in arch:
            traffic = ncu_traffic(f)
            if traffic is not None:
                tsrc = f'ARCHIVED ncu --set full capture profiles/{f} (largest launch of the class), not measured by this run'
                break
        step_tf = FLOPS_PER_ITER.get(size, 0) / (ms / args.steps * 1e-3) / 1e12
        roof = {'bound': d['bound'], 'kernel': f'{dom}: {d["what"]} (the class with the largest time per iteration)',
                'achieved': d['tflops'] if d['bound'] == 'tensor' else d['achieved_gbs'],
                'peak': pk['tflops'] if d['bound'] == 'tensor' else pk['hbm'],
                'unit': 'TFLOP/s' if d['bound'] == 'tensor' else 'GB/s', 'frac': d['frac'],
                'traffic': traffic, 'traffic_source': tsrc, 'peak_source': pk['src'],
                'timing': 'in-kernel %globaltimer stamps (dependencies satisfied -> last CTA done) of every tensor-core '
                          'launch over 3 profiled iterations run eagerly with their programmatic-dependent-launch overlap; '
                          'per-launch => burst peak',
                'classes': res, 'largest_launch': tops,
                'whole_step': {'tflops_all_convs': step_tf, 'frac_of_burst': step_tf / pk['tflops'] if pk['tflops'] else None,
                               'frac_of_sustained': step_tf / pk['tflops_sustained'] if pk['tflops_sustained'] else None,
                               'gflop_per_step': FLOPS_PER_ITER.get(size, 0) / 1e9}}

    # ---- configs[2] flavour: two independent images in flight on this GPU (separate nets / plans / streams):
    #      the latency-bound low-resolution levels of one image overlap the tensor-bound levels of the other ----
    conc = None
    if args.concurrent > 1 and rank == 0:
        conc = run_concurrent(args, dev, args.concurrent)

    # ---- end to end through the public call surface with host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, rank, world, dev, net, ds, lr_img, z_saved, barrier)
        if world == 1:
            e2e['pipelined_readback_variant'] = run_e2e(args, rank, world, dev, net, ds, lr_img, z_saved, barrier, True)

    if rank != 0:
        return
    its = world * args.steps / (ms * 1e-3)
    line = {
        'metric': 'DIP iters/s (512^2 4x SR)', 'value': its, 'unit': 'it/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f16 operands / f32 accumulate (tcgen05 kind::f16), f32 master weights + Adam',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD.format(s=size), 'size': size, 'factor': FACTOR,
                   'images': f'{world} independent image(s), one per GPU, no collective',
                   'noise': 'device Philox (value) / pre-drawn pinned host ring (e2e)',
                   'l2': 'working set per iteration ~1 GB >> 126 MB L2: no flush needed'},
        'roofline': roof, 'e2e': e2e, 'gpu_launches': launches_per_step * args.steps, 'clocks': clocks,
        'loss_first_last': [loss_first, loss_last], 'concurrent_images': conc,
        'conv_tflops_per_gpu': FLOPS_PER_ITER.get(size, 0) * (args.steps / (ms * 1e-3)) / 1e12,
    }
    if world == 1 and not args.no_cpu:
        # the reference's own modules on this box: host cores (the CPU baseline) and stock PyTorch CUDA eager (the
        # like-for-like GPU comparator); both in child processes (same module names as the drop-in)
        torch.cuda.synchronize()
        r = reference_subprocess(size, 'cpu', 6, 2)
        line['cpu_baseline'] = r.get('cpu_baseline', r)
        g = reference_subprocess(size, 'cuda', 30, 5)
        g2 = reference_subprocess(size, 'cuda', 30, 5, predrawn=True)
        line['gpu_eager_baseline'] = ({'value': g['value'], 'unit': 'it/s', 'sample': g['cpu_baseline']['sample'],
                                       'kind': g['cpu_baseline']['kind'],
                                       'value_predrawn_noise': g2.get('value'),
                                       'note': 'value: exactly the closure of DIP.py (8 M normals drawn on the host per '
                                               'step); value_predrawn_noise: inputs drawn ahead in pinned host memory, as '
                                               'in our e2e arm'} if 'value' in g else g)
        if not args.no_secondary:
            line['secondary'] = secondary_gan_eval(dev)
            line['secondary_gan_train'] = secondary_gan_train(dev)
    print(json.dumps(line), flush=True)


def run_concurrent(args, dev, n_img):
    """n_img independent images optimised concurrently on one GPU (each with its own net, plan and stream), as the
    64-image configuration (BASELINE configs[2]) would be scheduled.  Returns aggregate iterations / s."""
    import ctypes as C
    import dsr_b200
    from dsr_b200._lib import lib, check, StepBuffers
    size = args.size
    jobs = []
    for i in range(n_img):
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            hr = synthetic_pair(100 + i, size)
            ds = dsr_b200.Downsampler(3, FACTOR, 'lanczos2', phase=0.5, preserve_size=True)
            lr_img = ds(hr.unsqueeze(0).to(dev))[0].contiguous()
            torch.manual_seed(100 + i)
            net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                                   upsample_mode='bilinear').to(dev)
            z_saved = dsr_b200.get_noise(32, 'noise', (size, size)).to(dev).contiguous()
            z = z_saved.clone()
            net(z)
            net.zero_grad()
            plan = net._plans[(size, size)]
            tables = ds._tables_for(size, size, dev)
            oh, ow = ds.out_size(size, size)
            flat, gflat = net.flat_buffers()
            f32 = dict(dtype=torch.float32, device=dev)
            bufs = dict(m=torch.zeros_like(flat), v=torch.zeros_like(flat), out_hr=torch.empty((1, 3, size, size), **f32),
                        out_lr=torch.empty((3, oh, ow), **f32), g_lr=torch.empty((3, oh, ow), **f32),
                        g_hr=torch.empty((1, 3, size, size), **f32),
                        losses=torch.zeros(args.warmup + args.steps + 8, **f32))
            b = StepBuffers(flat.data_ptr(), gflat.data_ptr(), bufs['m'].data_ptr(), bufs['v'].data_ptr(),
                            net._bnflat.data_ptr(), z_saved.data_ptr(), z.data_ptr(), lr_img.data_ptr(),
                            bufs['out_hr'].data_ptr(), bufs['out_lr'].data_ptr(), bufs['g_lr'].data_ptr(),
                            bufs['g_hr'].data_ptr(), bufs['losses'].data_ptr())
        jobs.append(dict(st=st, net=net, ds=ds, plan=plan, tables=tables, b=b, keep=(bufs, z, z_saved, lr_img), t=0))

    def run(n):
        for j in jobs:
            check(lib.dsr_dip_run(j['plan'].handle, j['tables'].handle, C.byref(j['b']), LR_RATE, SIGMA, 77, j['t'] + 1, n,
                                  j['st'].cuda_stream), 'dsr_dip_run')
            j['t'] += n

    run(args.warmup)
    torch.cuda.synchronize()
    t0 = time.time()
    run(args.steps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    return {'images': n_img, 'value': n_img * args.steps / dt, 'unit': 'it/s (aggregate over the concurrent images)',
            'ms_per_round': dt / args.steps * 1e3}


def run_e2e(args, rank, world, dev, net, ds, lr_img, z_saved, barrier, pipelined=False):
    """Closure path of DIP.py:47-95 through get_params / optimize with HOST buffers for the step's input and
    results.  A ring of pre-drawn perturbed inputs lives in pinned host memory (drawing 8M normals per step on
    the CPU is the reference's 71 ms host cost and is not part of this metric); the copy of step t+1's input is
    enqueued on a side stream while step t computes."""
    import dsr_b200
    H, W = z_saved.shape[2], z_saved.shape[3]
    ring_n = 4
    zs = z_saved.cpu()
    ring = [(zs + torch.randn(zs.shape) * SIGMA).pin_memory() for _ in range(ring_n)]
    dev_z = [torch.empty_like(z_saved) for _ in range(2)]
    host_hr = torch.empty((1, 3, H, W), dtype=torch.float32).pin_memory()
    host_lr = torch.empty((1, 3, H // FACTOR, W // FACTOR), dtype=torch.float32).pin_memory()
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    fwd_done = torch.cuda.Event()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    target = lr_img.unsqueeze(0)
    mse = torch.nn.MSELoss()
    state = {'i': 0}

    def prefetch(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            dev_z[slot].copy_(ring[i % ring_n], non_blocking=True)
            ready[slot].record(copy_stream)

    def closure():
        i = state['i']
        slot = i % 2
        torch.cuda.current_stream().wait_event(ready[slot])
        out_hr = net(dev_z[slot])
        consumed[slot].record()
        prefetch(i + 2)
        out_lr = ds(out_hr)
        loss = mse(out_lr, target)
        if pipelined:
            # variant: the three results leave over PCIe on a side stream WHILE the backward pass runs; the host blocks
            # on their arrival only
            fwd_done.record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(fwd_done)
                host_hr.copy_(out_hr.detach(), non_blocking=True)
                host_lr.copy_(out_lr.detach(), non_blocking=True)
                host_loss.copy_(loss.detach(), non_blocking=True)
            loss.backward()
            d2h_stream.synchronize()
        else:
            # reference order (DIP.py:68 then :90-91): backward first, then the blocking reads on the compute stream
            loss.backward()
            host_hr.copy_(out_hr.detach(), non_blocking=True)
            host_lr.copy_(out_lr.detach(), non_blocking=True)
            host_loss.copy_(loss.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        state['i'] += 1
        return loss

    for s in range(2):
        consumed[s].record()
    prefetch(0)
    prefetch(1)
    params = dsr_b200.get_params('net', net, z_saved)
    dsr_b200.optimize('adam', params, closure, LR_RATE, max(args.warmup, 3))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    dsr_b200.optimize('adam', params, closure, LR_RATE, args.steps)
    e1.record()
    barrier()
    wall = time.time() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    h2d = z_saved.numel() * 4
    d2h = host_hr.numel() * 4 + host_lr.numel() * 4 + 4
    return {'value': world * args.steps / (ms * 1e-3), 'unit': 'it/s', 'h2d_bytes_per_step': h2d,
            'd2h_bytes_per_step': d2h, 'ms_per_step': ms / args.steps,
            'path': 'get_net/Downsampler/get_params/optimize + DIP.py-style closure; z copied from a pre-drawn ring of '
                    'perturbed inputs in pinned host memory (the reference draws 8 M normals on the CPU per step: 71 ms, '
                    'not part of this metric); out_HR/out_LR/loss read back each step ' +
                    ('on a side stream during the backward pass, the host blocks on their arrival (pipelined variant)'
                     if pipelined else 'AFTER the backward pass on the compute stream, the host blocks (reference order, '
                     'DIP.py:68,90-91)')}


# -------------------------------------------------------------------------------------------------
# secondary workload: SRResNet generator inference (BASELINE configs[3])
# -------------------------------------------------------------------------------------------------
GAN_FLOPS_96 = 98.131378176e9        # per 96 x 96 image at x8 (oracle.gan_oracle.flops_per_image)


def gan_cpu_images_per_s(n_img, budget_s):
    """The CPU restatement of the reference generator (oracle/gan_oracle.py, eval mode, fp32, all host threads) on a
    bounded sample: n_img single-image forwards at 96 x 96 (eval_GAN.py uses batch size 1, :81)."""
    from oracle import gan_oracle as G
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = G.init_state_dict(8)
    times = []
    t_begin = time.time()
    with torch.no_grad():
        for i in range(n_img + 1):
            x = torch.rand(1, 3, 96, 96)
            t0 = time.time()
            G.generator_forward(sd, x, 8)
            if i > 0:
                times.append(time.time() - t0)
            if time.time() - t_begin > budget_s and len(times) >= 2:
                break
    times.sort()
    return 1.0 / times[len(times) // 2], len(times), cores


def run_gan_reference(args, rank, world):
    if rank != 0:
        return
    ips, n, cores = gan_cpu_images_per_s(max(args.steps, 3), budget_s=120.0)
    line = {'impl': 'reference', 'metric': 'SRGAN generator images/s (96x96 LR patches, x8)', 'value': ips,
            'unit': 'images/s', 'n_gpus': args.gpus, 'steps': n, 'warmup': 1, 'ms_per_step': 1000.0 / ips,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'SRResNet generator inference, 96x96 LR patches -> 768x768 (BASELINE configs[3]); '
                                   'one image per step on the host cores', 'factor': 8},
            'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{n} single-image forwards (median), oracle/gan_oracle.py on torch CPU'},
            'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def run_gan(args, rank, local_rank, world):
    """One step = Generator(factor 8).eval() on a batch of `--batch` (256) LR patches of 96 x 96 -> 768 x 768."""
    import torch.distributed as dist
    import dsr_b200
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    B, hw = args.batch, 96

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(rank)
    g = dsr_b200.Generator(8).to(dev).eval()
    x = torch.rand(B, 3, hw, hw, device=dev)
    for _ in range(args.warmup):
        y = g(x)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        y = g(x)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    # ---- end to end: LR batch from pinned host memory, SR batch read back to pinned host memory ----
    hx = torch.rand(B, 3, hw, hw).pin_memory()
    hy = torch.empty((B, 3, 8 * hw, 8 * hw), dtype=torch.float32).pin_memory()
    chunk = g.max_chunk
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()

    def e2e_step():
        # the caller pipelines chunks of the batch: chunk i's SR images leave over PCIe while chunk i+1 is computed
        for b0 in range(0, B, chunk):
            yc = g(hx[b0:b0 + chunk].to(dev, non_blocking=True))
            done = torch.cuda.Event()
            done.record(main)
            yc.record_stream(copy_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                hy[b0:b0 + chunk].copy_(yc, non_blocking=True)
        main.wait_stream(copy_stream)

    for _ in range(2):
        e2e_step()
    barrier()
    n_e2e = max(2, min(args.steps, 5))
    t0 = time.time()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(n_e2e):
        e2e_step()
    f1.record()
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), (time.time() - t0) * 1e3)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return
    from dsr_b200._lib import lib
    plan = next(iter(g._plans.values()))
    launches = lib.dsr_gen_last_launches(plan.handle) * ((B + g.max_chunk - 1) // g.max_chunk)
    pk = peaks()
    ips = world * args.steps * B / (ms * 1e-3)
    tfl = GAN_FLOPS_96 * B * args.steps / (ms * 1e-3) / 1e12
    line = {'metric': 'SRGAN generator images/s (96x96 LR patches, x8, batch 256)', 'value': ips, 'unit': 'images/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16 operands / f32 accumulate (tcgen05 kind::f16), eval-mode BatchNorm folded', 'data': 'synthetic',
            'config': {'workload': f'SRResNet generator inference (eval_GAN), batch {B} of 96x96 LR patches -> 768x768 '
                                   f'per GPU (BASELINE configs[3]; the reference Generator builds x8 / x16 only), '
                                   f'chunks of {g.max_chunk} images per library call',
                       'batch': B, 'factor': 8, 'l2': 'activations of a 32-image chunk are 0.3-2.4 GB >> 126 MB L2'},
            'roofline': {'bound': 'tensor', 'kernel': 'whole forward: conv_halo2_kernel<1> (3x3 64->64 / 64->256+shuffle) '
                         'conv9_out_kernel (9x9 64->3 + tanh, kx taps folded into N) and gen_conv1_kernel (9x9 3->64, CUDA cores)', 'achieved': tfl, 'peak': pk['tflops'], 'unit': 'TFLOP/s',
                         'frac': tfl / pk['tflops'] if pk['tflops'] else None, 'traffic': None, 'peak_source': pk['src'],
                         'gflop_per_image': GAN_FLOPS_96 / 1e9},
            'e2e': {'value': world * n_e2e * B / (ms_e2e * 1e-3), 'unit': 'images/s', 'h2d_bytes_per_step': hx.numel() * 4,
                    'd2h_bytes_per_step': hy.numel() * 4, 'ms_per_step': ms_e2e / n_e2e,
                    'path': 'dsr_b200.Generator.__call__ per 32-image chunk of a pinned host batch; each chunk\'s SR images '
                            'are copied back to pinned host memory on a second stream while the next chunk is computed'},
            'gpu_launches': launches * args.steps, 'clocks': clocks}
    if world == 1 and not args.no_cpu:
        ips_cpu, n, cores = gan_cpu_images_per_s(8, budget_s=30.0)
        line['cpu_baseline'] = {'value': ips_cpu, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                                'sample': f'{n} single-image forwards at 96x96 (median), oracle/gan_oracle.py on torch CPU'}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# SRGAN training step (BASELINE configs[4]): do_epoch of train_GAN.py:38-71
# -------------------------------------------------------------------------------------------------
GT_B, GT_LR = 8, 24                       # train_GAN.py:168 batch size, :269 HR patch 192 = 8 x 24
# algorithmic conv FLOPs per patch: generator 6.133 GF (98.13 / 16), discriminator 7.07 GF, VGG19 features[:36]
# at 224 x 224 39.01 GF; one step = G fwd + bwd (3x), D 3 fwd + 2 bwd (7x), VGG 2 fwd + 1 data-gradient (3x)
GT_FLOPS_PER_PATCH = (3 * 6.1332 + 7 * 7.07 + 3 * 39.01) * 1e9


def _gan_train_objects(dev, seed):
    import dsr_b200
    from dsr_b200 import gan_train as GT
    torch.manual_seed(seed)
    G = dsr_b200.Generator(8).train()
    D = GT.Discriminator((GT_LR * 8, GT_LR * 8)).train()
    torch.manual_seed(seed + 1000)
    V = GT.Vgg19Loss(pretrained=False).to(dev)          # no pretrained file offline: random VGG19
    return G, D, V, GT


def run_gan_train_reference(args, rank, world):
    """The reference's own do_epoch (train_GAN.GAN_ISR_train of the unmodified checkout in baseline/_ref, its own
    modules, torch CPU fp32, all host threads; `--ref-device cuda`: stock CUDA eager, TF32 off) on one batch of 8."""
    if rank != 0:
        return
    import subprocess
    steps = max(2, min(args.steps, 3)) if args.ref_device == 'cpu' else max(args.steps, 3)
    cmd = [sys.executable, os.path.join(ROOT, 'tools', 'run_reference_gan.py'), '--impl', 'reference', '--device',
           args.ref_device, '--batch', str(GT_B), '--lr-size', str(GT_LR), '--epochs', str(steps)]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout.strip().splitlines()
    r = json.loads(out[-1]) if out else {'unavailable': 'tools/run_reference_gan.py produced no output'}
    if 'unavailable' in r:
        print(json.dumps({'impl': 'reference', 'unavailable': r['unavailable']}), flush=True)
        return
    pps = GT_B * r['steps_per_s']
    cores = os.cpu_count() or 1
    line = {'impl': 'reference', 'metric': 'SRGAN training patches/s (do_epoch: G, D, VGG19 loss, two Adam steps)',
            'value': pps, 'unit': 'patches/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': 0,
            'ms_per_step': 1000.0 / r['steps_per_s'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'SRGAN training step (train_GAN.do_epoch), batch {GT_B} of {GT_LR}x{GT_LR} -> 192x192 '
                                   f'patches per GPU, random-weight VGG19 (BASELINE configs[4])', 'batch': GT_B, 'factor': 8},
            'cpu_baseline': {'value': pps, 'unit': 'patches/s', 'cores': cores if args.ref_device == 'cpu' else 0,
                             'kind': 'reference', 'device': args.ref_device,
                             'sample': f'{steps} do_epoch calls of the unmodified reference (incl. module construction and '
                                       f'one logging pass), {r["seconds"]:.1f} s'},
            'e2e': {'value': pps, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'losses': [r['loss_D'], r['loss_G']], 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def gan_train_eager_baseline():
    """Like-for-like GPU comparator (harness only): the UNMODIFIED reference's own loop over its own modules on this GPU
    in stock PyTorch CUDA eager (fp32, TF32 off, cuDNN), tools/run_reference_gan.py --impl reference.  Steady-state rate
    = the difference of a 13-step and a 3-step run (module construction, cuDNN start-up and the one logging pass cancel)."""
    import subprocess

    def run(epochs):
        cmd = [sys.executable, os.path.join(ROOT, 'tools', 'run_reference_gan.py'), '--impl', 'reference', '--device', 'cuda',
               '--batch', str(GT_B), '--lr-size', str(GT_LR), '--epochs', str(epochs)]
        out = subprocess.run(cmd, capture_output=True, text=True).stdout.strip().splitlines()
        return json.loads(out[-1]) if out else {'unavailable': 'no output'}
    try:
        a, b = run(3), run(13)
        if 'unavailable' in a or 'unavailable' in b:
            return {'unavailable': a.get('unavailable') or b.get('unavailable')}
        sps = 10.0 / max(b['seconds'] - a['seconds'], 1e-6)
        return {'value': GT_B * sps, 'unit': 'patches/s', 'ms_per_step': 1000.0 / sps, 'kind': 'reference',
                'what': 'unmodified train_GAN.GAN_ISR_train over the reference modules, stock PyTorch CUDA eager fp32 (TF32 '
                        'off), same batch; (13-step run - 3-step run) / 10'}
    except Exception as e:
        return {'unavailable': repr(e)[:200]}


def run_gan_train(args, rank, local_rank, world):
    """One step = GanTrainStep.do_epoch on this rank's batch of 8 patches; world > 1: data-parallel replicas, NCCL
    all-reduce (mean) of the flat discriminator (321 MB) and generator (6.8 MB) gradients."""
    import torch.distributed as dist
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    from oracle import gan_train_oracle as O          # harness only: the synthetic LR / HR batch recipe
    G, D, V, GT = _gan_train_objects(dev, 0)
    step = GT.GanTrainStep(G, D, V, 1e-4, GT_B, (GT_LR, GT_LR), dev, data_parallel=world > 1)
    LR, HR = O.synthetic_batch(100 + rank, GT_B, (GT_LR, GT_LR), 8)          # every rank its own shard of the global batch
    dLR, dHR = LR.to(dev), HR.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step.do_epoch(dLR, dHR)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lD, lG = step.do_epoch(dLR, dHR)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    # ---- end to end: the batch comes from pinned host memory, both losses are read back every step (train_GAN.py:99-100)
    hLR, hHR = LR.pin_memory(), HR.pin_memory()
    for _ in range(2):
        a, b = step.do_epoch(hLR, hHR)
        a.item(), b.item()
    barrier()
    n_e2e = max(3, min(args.steps, 10))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    f0.record()
    for _ in range(n_e2e):
        a, b = step.do_epoch(hLR, hHR)
        losses = (a.item(), b.item())
    f1.record()
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), (time.time() - t0) * 1e3)
    # ---- the collective alone: mean all-reduce of the two flat gradient buffers
    ar = None
    if world > 1:
        x = step.xch
        for _ in range(2):
            x.allreduce_mean(step.fd.gflat)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            x.allreduce_mean(step.fd.gflat)
            x.allreduce_mean(step.fg.gflat)
        g1.record()
        barrier()
        ar_ms = g0.elapsed_time(g1) / 5
        nbytes = (step.fd.gflat.numel() + step.fg.gflat.numel()) * 4
        t = torch.tensor([ms, ms_e2e, ar_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ar_ms = float(t[0]), float(t[1]), float(t[2])
        ar = {'ms_per_step': ar_ms, 'bytes_per_step': nbytes, 'algbw_GBps': nbytes / (ar_ms * 1e-3) / 1e9,
              'busbw_GBps': 2 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9,
              'note': 'timed alone; inside the step the discriminator all-reduce sits in the discriminator chain (second '
                      'stream), which runs beside the generator chain (VGG loss, generator backward)'}
    if rank != 0:
        return
    pk = peaks()
    pps = world * GT_B * args.steps / (ms * 1e-3)
    tfl = GT_FLOPS_PER_PATCH * GT_B * args.steps / (ms * 1e-3) / 1e12
    line = {'metric': 'SRGAN training patches/s (do_epoch: G, D, VGG19 loss, two Adam steps)', 'value': pps,
            'unit': 'patches/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16 operands / f32 accumulate (tcgen05 kind::f16), f32 master weights and Adam', 'data': 'synthetic',
            'config': {'workload': f'SRGAN training step (train_GAN.do_epoch), batch {GT_B} of {GT_LR}x{GT_LR} -> 192x192 '
                                   f'patches per GPU, random-weight VGG19 (BASELINE configs[4])', 'batch': GT_B, 'factor': 8,
                       'parallelism': f'dp{world}' if world > 1 else 'single',
                       'streams': 'discriminator chain, generator chain and the VGG pass of the real batch on three '
                                  'streams of one process (they share only the generated batch)',
                       'l2': 'activations of one step are ~1.4 GB >> 126 MB L2'},
            'roofline': {'bound': 'tensor', 'kernel': 'whole step: gconv_kernel / gwgrad_kernel (bf16 implicit GEMM) + element-wise family',
                         'achieved': tfl, 'peak': pk['tflops'], 'unit': 'TFLOP/s',
                         'frac': tfl / pk['tflops'] if pk['tflops'] else None, 'traffic': None, 'peak_source': pk['src'],
                         'gflop_per_patch': GT_FLOPS_PER_PATCH / 1e9},
            'e2e': {'value': world * GT_B * n_e2e / (ms_e2e * 1e-3), 'unit': 'patches/s',
                    'h2d_bytes_per_step': (LR.numel() + HR.numel()) * 4, 'd2h_bytes_per_step': 8,
                    'ms_per_step': ms_e2e / n_e2e,
                    'path': 'GanTrainStep.do_epoch(LR, HR) with the batch in pinned host memory; both losses read back with '
                            '.item() every step (train_GAN.py:99-100)'},
            'losses': list(losses), 'allreduce': ar,
            'gpu_launches': None, 'clocks': clocks}
    line['gpu_launches'] = step.launches_per_step * args.steps
    if world == 1 and not args.no_cpu:
        line['gpu_eager_baseline'] = gan_train_eager_baseline()
    print(json.dumps(line), flush=True)



def secondary_gan_train(dev):
    """BASELINE configs[4] beside the headline: a short single-GPU run of the fused SRGAN training step (batch 8 of
    24x24 -> 192x192 patches, random-weight VGG19); the full line incl. data parallelism is `--workload gan_train`."""
    try:
        from oracle import gan_train_oracle as O          # harness only: the synthetic LR / HR batch recipe
        G, D, V, GT = _gan_train_objects(dev, 0)
        step = GT.GanTrainStep(G, D, V, 1e-4, GT_B, (GT_LR, GT_LR), dev)
        LR, HR = O.synthetic_batch(100, GT_B, (GT_LR, GT_LR), 8)
        LR, HR = LR.to(dev), HR.to(dev)
        steps = 10
        for _ in range(3):
            step.do_epoch(LR, HR)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step.do_epoch(LR, HR)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        pk = peaks()
        tf = GT_FLOPS_PER_PATCH * GT_B / (ms * 1e-3) / 1e12
        return {'workload': f'SRGAN training step (train_GAN.do_epoch), batch {GT_B} of {GT_LR}x{GT_LR} -> 192x192 patches, '
                            f'random-weight VGG19 (BASELINE configs[4]), {steps} steps', 'value': GT_B / (ms * 1e-3),
                'unit': 'patches/s', 'ms_per_step': ms, 'tflops': tf,
                'frac_of_burst_bf16': tf / pk['tflops'] if pk['tflops'] else None,
                'gpu_launches_per_step': step.launches_per_step}
    except Exception as e:                                 # never let the secondary figure take the headline down
        return {'workload': 'SRGAN training step', 'error': repr(e)[:200]}


def secondary_gan_eval(dev):
    """BASELINE configs[3] beside the headline: a short SRResNet generator inference run (batch 64 of 96x96 LR patches,
    x8) through dsr_b200.Generator; the full batch-256 line is `--workload gan_eval`."""
    import dsr_b200
    B, hw, steps = 64, 96, 3
    torch.manual_seed(0)
    g = dsr_b200.Generator(8).to(dev).eval()
    x = torch.rand(B, 3, hw, hw, device=dev)
    for _ in range(2):
        g(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ips = B / (ms * 1e-3)
    pk = peaks()
    tf = GAN_FLOPS_96 * ips / 1e12
    return {'workload': f'SRResNet generator inference (eval_GAN), batch {B} of 96x96 LR patches -> 768x768 (BASELINE '
                        f'configs[3]), {steps} steps', 'value': ips, 'unit': 'images/s', 'ms_per_step': ms,
            'tflops': tf, 'frac_of_burst_bf16': tf / pk['tflops'] if pk['tflops'] else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='dip', choices=['dip', 'gan_eval', 'gan_train'],
                    help='dip: the headline DIP iteration (BASELINE configs[1]); gan_eval: SRResNet generator inference (configs[3]); '
                         'gan_train: SRGAN training step, data-parallel over the ranks (configs[4])')
    ap.add_argument('--batch', type=int, default=256, help='gan_eval: LR patches per step')
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--concurrent', type=int, default=2, help='also time this many images in flight on the GPU')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-secondary', action='store_true')
    ap.add_argument('--ref-predrawn', action='store_true', help='--impl reference: perturbed inputs drawn ahead of the loop')
    ap.add_argument('--ref-device', default='cpu', choices=['cpu', 'cuda'],
                    help='--impl reference: host cores (the CPU baseline) or stock PyTorch CUDA eager (harness only)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.steps is None:
        args.steps = {'dip': 200, 'gan_eval': 10 if args.impl == 'ours' else 8, 'gan_train': 20}[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        {'dip': run_reference, 'gan_eval': run_gan_reference, 'gan_train': run_gan_train_reference}[args.workload](args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)')
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        {'dip': run_ours, 'gan_eval': run_gan, 'gan_train': run_gan_train}[args.workload](args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
