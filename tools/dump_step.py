"""GPU-side helper: run the teacher-forced step of a reference fixture through the CUDA path and save everything the
CPU-side analysis needs (output, loss, all gradients, raw conv outputs) to gpurun_out/cuda_<fixture>.pt; also reports
run-to-run agreement of two fresh networks in this process (bit-identical with DSR_DETERMINISTIC=1).

    python tools/dump_step.py step_64x64.pt [step_72x88.pt ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
sys.path.insert(0, ROOT)


def run(fx, keep_taps):
    import dsr_b200
    torch.manual_seed(fx['seed'])
    net = dsr_b200.get_net(32, 'skip', fx.get('pad', 'reflection'), skip_n33d=128, skip_n33u=128, skip_n11=4,
                           num_scales=5, upsample_mode=fx.get('upsample_mode', 'bilinear'))
    if 'z0' in fx:
        z0 = fx['z0']
    else:        # large fixtures keep the seed only (oracle/make_golden_large.py): same CPU draws as the reference
        ni = torch.zeros(1, 32, fx['H'], fx['W']).uniform_() * 0.1
        ds = dsr_b200.Downsampler(3, fx['factor'], 'lanczos2', phase=0.5, preserve_size=True)   # draws like the reference's
        z0 = ni + ni.clone().normal_() * fx['reg_noise_std']
    net = net.cuda()
    ds = dsr_b200.Downsampler(3, fx['factor'], 'lanczos2', phase=0.5, preserve_size=True).cuda()
    z = z0.cuda()
    if os.environ.get('DSR_DUMP_CHECKER'):     # CUDA-core checker kernels (sequential fp32 FMA) instead of tcgen05
        net(z)
        net.set_debug_conv(True)
    out = net(z)
    out_lr = ds(out)
    loss = torch.nn.MSELoss()(out_lr, fx['lr_img'].cuda())
    loss.backward()
    torch.cuda.synchronize()
    res = {'out_hr': out.detach().cpu(), 'out_lr': out_lr.detach().cpu(), 'loss': float(loss),
           'grads': {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()}}
    if keep_taps:
        hw = (fx['H'], fx['W'])
        taps = {}
        for i in range(5):
            for tag in ('d1', 'd2', 'u1', 'u2'):
                for suf in ('_raw', '_dr'):
                    taps[f'L{i}.{tag}{suf}'] = net.debug_tensor(f'L{i}.{tag}{suf}', hw).cpu()
            taps[f'L{i}.cat'] = net.debug_tensor(f'L{i}.cat', hw).cpu()
            taps[f'L{i}.sraw'] = net.debug_tensor(f'L{i}.sraw', hw).cpu()
        taps['gscale'] = net.debug_tensor('gscale', hw).cpu()
        res['taps'] = taps
    return res


def main():
    out_dir = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    for name in sys.argv[1:]:
        fx = torch.load(os.path.join(ROOT, 'tests', 'golden', name))
        a = run(fx, fx['H'] <= 128)
        b = run(fx, False)
        same_out = torch.equal(a['out_hr'], b['out_hr'])
        ga = torch.cat([v.flatten() for v in a['grads'].values()]).double()
        gb = torch.cat([v.flatten() for v in b['grads'].values()]).double()
        cos = float(ga @ gb / (ga.norm() * gb.norm()))
        print(f'{name}: det={os.environ.get("DSR_DETERMINISTIC", "0")} loss {a["loss"]:.8f} / {b["loss"]:.8f} '
              f'(fixture {fx["losses"][0] if "losses" in fx else fx["loss"]:.8f})  out bit-equal {same_out}  grads bit-equal {torch.equal(ga, gb)} '
              f'cos {cos:.9f}  max|dg| {float((ga - gb).abs().max()):.3e}', flush=True)
        rel = float((a['out_hr'] - fx['out_hr']).norm() / fx['out_hr'].norm())
        print(f'   out rel L2 vs reference {rel:.3e}', flush=True)
        tag = 'chk' if os.environ.get('DSR_DUMP_CHECKER') else 'det' if os.environ.get('DSR_DETERMINISTIC') else 'def'
        torch.save(a, os.path.join(out_dir, f'cuda_{tag}_{name}'))


if __name__ == '__main__':
    main()
