"""GPU diagnostic (not collected by pytest): localises a numerical fault to one kernel.

    python tests/diag_gpu.py --mode checker     # elementwise kernels + orchestration vs the oracle (CUDA-core convs)
    python tests/diag_gpu.py --mode replay      # every tcgen05 launch vs its checker kernel, same inputs
    python tests/diag_gpu.py --mode tc          # whole step with the tcgen05 kernels vs the oracle

Prints one line per tensor: relative L2 error against the CPU oracle (oracle/dip_oracle.py).
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
sys.path.insert(0, ROOT)

import dsr_b200  # noqa: E402
from dsr_b200._lib import lib, check  # noqa: E402
from oracle import dip_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def nchw(t, padded):
    """plan tensor [H(+2)][W(+2)][C] -> [1,C,H,W] interior, fp32 on the CPU"""
    if padded:
        t = t[1:-1, 1:-1]
    return t.float().permute(2, 0, 1).unsqueeze(0).cpu()


def unperm_cat(t):  # packed concat channels -> reference order (skip first)
    return torch.cat([t[:, 128:132], t[:, 0:128]], dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='checker', choices=['checker', 'replay', 'tc'])
    ap.add_argument('--fixture', default='step_64x64.pt')
    ap.add_argument('--only', default='', help='replay mode: what:layer, e.g. 2:L0.u2')
    ap.add_argument('--size', type=int, default=0, help='use a random problem of this size instead of the fixture')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    if args.size:
        H = W = args.size
        torch.manual_seed(5)
        net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                               upsample_mode='bilinear')
        g = torch.Generator().manual_seed(11)
        z0 = torch.rand(1, 32, H, W, generator=g) * 0.1 + torch.randn(1, 32, H, W, generator=g) * 0.05
        hr = torch.rand(1, 3, H, W, generator=g)
        lr_img = O.downsample(hr, 4)
        factor = 4
    else:
        fx = torch.load(os.path.join(ROOT, 'tests', 'golden', args.fixture))
        H, W, factor = fx['H'], fx['W'], fx['factor']
        torch.manual_seed(fx['seed'])
        net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                               upsample_mode='bilinear')
        z0, lr_img = fx['z0'], fx['lr_img']
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}

    # ---- oracle ----
    taps = {}
    loss_o, out_o, grads_o = O.step_loss_and_grads(sd, z0, lr_img, factor, taps=taps)
    print(f'oracle loss {float(loss_o):.6f}')
    if not args.size:
        print('oracle vs reference fixture: out', rel(out_o, fx['out_hr']), 'loss', abs(float(loss_o) - fx['losses'][0]))

    # ---- ours ----
    net = net.to(dev)
    ds = dsr_b200.Downsampler(3, factor, 'lanczos2', phase=0.5, preserve_size=True)
    zc = z0.to(dev)
    hw = (H, W)

    def run(debug):
        net.zero_grad()
        if hw in net._plans:
            net.set_debug_conv(debug)
        else:
            out = net(zc)          # creates the plan
            net.set_debug_conv(debug)
        out = net(zc)
        out_lr = ds(out)
        loss = torch.nn.functional.mse_loss(out_lr, lr_img.to(dev))
        loss.backward()
        torch.cuda.synchronize()
        return out.detach(), out_lr.detach(), float(loss)

    def report(tag):
        print(f'==== {tag}: forward intermediates (rel L2 vs oracle) ====')
        for i in range(5):
            n = O.layer_names(i)
            P = f'L{i}.'
            def bias(name):
                return sd[n[name] + '.bias'].view(1, -1, 1, 1)
            rows = [
                ('sraw', nchw(net.debug_tensor(P + 'sraw', hw), 0) + bias('skip_conv'), taps[P + 'skip_raw']),
                ('d1_raw', nchw(net.debug_tensor(P + 'd1_raw', hw), 0) + bias('d1_conv'), taps[P + 'd1_raw']),
                ('d2_raw', nchw(net.debug_tensor(P + 'd2_raw', hw), 0) + bias('d2_conv'), taps[P + 'd2_raw']),
                ('d2_act', nchw(net.debug_tensor(P + 'd2_act', hw), 1), taps[P + 'x_next']),
                ('cat', unperm_cat(nchw(net.debug_tensor(P + 'cat', hw), 1)[:, :132]), taps[P + 'cat']),
                ('u1_raw', nchw(net.debug_tensor(P + 'u1_raw', hw), 0) + bias('u1_conv'), taps[P + 'u1_raw']),
                ('u2_raw', nchw(net.debug_tensor(P + 'u2_raw', hw), 0) + bias('u2_conv'), taps[P + 'u2_raw']),
                ('u2_act', nchw(net.debug_tensor(P + 'u2_act', hw), 1), taps[P + 'out']),
            ]
            print(P, '  '.join(f'{k} {rel(a, b.detach()):.2e}' for k, a, b in rows))
        print(f'==== {tag}: backward intermediates ====')
        gsc = net.debug_tensor('gscale', hw).flatten().cpu()
        S = float(gsc[6])
        print(f'gradient scale used {S}, amax {float(gsc[4]):.3e}, non-finite passes {float(gsc[5])}')
        for i in range(5):
            P = f'L{i}.'
            rows = []
            for ours, theirs in (('u2_dr', 'u2_raw'), ('u1_dr', 'u1_raw'), ('d2_dr', 'd2_raw'), ('d1_dr', 'd1_raw')):
                g = taps[P + theirs].grad
                rows.append((ours, nchw(net.debug_tensor(P + ours, hw), 1) / S, g))
            gc = taps[P + 'cat'].grad
            rows.append(('dsraw', nchw(net.debug_tensor(P + 'dsraw', hw), 0) / S, taps[P + 'skip_raw'].grad))
            print(P, '  '.join(f'{k} {rel(a, b):.2e}' for k, a, b in rows), f'|gcat| {float(gc.norm()):.2e}')

    def report_grads(tag):
        print(f'==== {tag}: parameter gradients (rel L2 / cosine vs oracle; dead parameters skipped) ====')
        dead = set(O.dead_param_keys())
        worst = 0.0
        for (name, p) in net.named_parameters():
            if name in dead:
                continue
            go = grads_o[name]
            g = p.grad.detach().cpu()
            r = rel(g, go)
            cos = float((g.double().flatten() @ go.double().flatten()) / max(float(g.double().norm() * go.double().norm()), 1e-30))
            worst = max(worst, 1 - cos)
            print(f'  {name:34s} rel {r:.2e} cos {cos:.6f} |g| {float(go.norm()):.2e}')
        print('worst 1-cos', worst)

    if args.mode == 'checker':
        out, out_lr, loss = run(1)
        print('launches (bwd)', net.last_launches(hw))
        print(f'[checker] loss {loss:.6f} vs oracle {float(loss_o):.6f}; out rel {rel(out, out_o):.3e}')
        report('checker')
        report_grads('checker')
    elif args.mode == 'tc':
        out, out_lr, loss = run(0)
        print(f'[tc] loss {loss:.6f} vs oracle {float(loss_o):.6f}; out rel {rel(out, out_o):.3e}')
        report('tc')
        report_grads('tc')
    else:
        out, out_lr, loss = run(1)      # consistent workspace from the checker path
        print(f'[checker] loss {loss:.6f} vs oracle {float(loss_o):.6f}; out rel {rel(out, out_o):.3e}')
        plan = net._plans[hw]
        stream = torch.cuda.current_stream().cuda_stream
        for what, suffix, wname in ((0, '_raw', 'fprop'), (2, '_dw', 'wgrad'), (1, '_gin', 'dgrad')):
            for i in range(5):
                for tag in ('d1', 'd2', 'u1', 'u2'):
                    layer = f'L{i}.{tag}'
                    if what == 1 and layer == 'L0.d1':
                        continue
                    if args.only and args.only != f'{what}:{layer}':
                        continue
                    res = {}
                    for chk in (1, 0):
                        check(lib.dsr_plan_debug_replay(plan.handle, layer.encode(), what, chk, stream), 'replay')
                        torch.cuda.synchronize()
                        if chk == 0:
                            import ctypes
                            code = ctypes.c_int()
                            check(lib.dsr_plan_device_error(plan.handle, ctypes.byref(code)))
                            if code.value:
                                print(f'!! device error word {code.value} after {wname} {layer}', flush=True)
                        res[chk] = net.debug_tensor(layer + suffix, hw).float().cpu()
                        if what == 0:
                            res[(chk, 's')] = net.debug_tensor(layer + '_stats', hw).float().cpu()
                    line = f'{wname:6s} {layer}: tc vs checker rel {rel(res[0], res[1]):.3e}  max|ref| {float(res[1].abs().max()):.3e}'
                    if what == 0:
                        line += f'  stats rel {rel(res[(0, "s")], res[(1, "s")]):.3e}'
                    print(line, flush=True)
        import ctypes
        code = ctypes.c_int()
        check(lib.dsr_plan_device_error(plan.handle, ctypes.byref(code)))
        print('device error word', code.value)


if __name__ == '__main__':
    main()
