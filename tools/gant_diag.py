"""Stage-by-stage check of the SRGAN training kernels against the oracle (run on the GPU box):

    python tools/gant_diag.py [batch lr_h lr_w]

Every stage prints its errors and the script carries on after a mismatch, so one run shows everything that is wrong."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
sys.path.insert(0, ROOT)

import dsr_b200                                   # noqa: E402
from dsr_b200 import gan_train as GT              # noqa: E402
from oracle import gan_train_oracle as O          # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


def nhwc_to_nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def grad_report(tag, named, flat_views, ref):
    worst = (2.0, None)
    allg, allr = [], []
    for (k, _), v in zip(named, flat_views):
        r = ref[k]
        if float(r.norm()) < 1e-12 and float(v.norm()) < 1e-12:
            continue
        c = cos(v, r)
        allg.append(v.flatten().double()); allr.append(r.flatten().double())
        if c < worst[0]:
            worst = (c, k)
        if c < 0.98 or abs(float(v.norm()) / (float(r.norm()) + 1e-30) - 1) > 0.05:
            print(f'   {tag} {k:48s} cos {c:.4f}  |g| {float(v.norm()):.3e} vs {float(r.norm()):.3e}')
    a, b = torch.cat(allg), torch.cat(allr)
    print(f'{tag}: whole-gradient cosine {cos(a, b):.5f}  norm ratio {float(a.norm() / b.norm()):.4f}  worst tensor {worst[1]} {worst[0]:.4f}')


def main():
    B, lh, lw = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 8, 8)
    f = 8
    H, W = lh * f, lw * f
    torch.manual_seed(3)
    G = dsr_b200.Generator(f).to(dev).train()
    D = GT.Discriminator((H, W)).to(dev).train()
    V = GT.Vgg19Loss(pretrained=False).to(dev)
    LR, HR = O.synthetic_batch(5, B, (lh, lw), f)
    LR, HR = LR.to(dev), HR.to(dev)
    sdG = {k: v.detach().clone() for k, v in G.state_dict().items()}
    sdD = {k: v.detach().clone() for k, v in D.state_dict().items()}
    sdV = {k: v.detach().clone() for k, v in V.state_dict().items()}

    # ---------------- discriminator forward / backward ----------------
    taps = {}
    d = O._leaf(sdD)
    p_ref = O.discriminator_train(d, HR, {}, taps)
    p = D(HR)
    tr = D._trainer_for(HR)[0]
    print('device error', tr.device_error())
    for k in ['d_h0'] + [f'd_h{i}' for i in range(1, 8)]:
        print(f'D fwd {k}: rel {rel(nhwc_to_nchw(tr.tensor(k)), taps[k]):.3e}')
    print('D prob', p.flatten().tolist(), 'ref', p_ref.flatten().tolist())
    loss = O.bce(p, 1.0)
    loss_ref = O.bce(p_ref, 1.0)
    keys = O.param_keys(sdD)
    gref = dict(zip(keys, torch.autograd.grad(loss_ref, [d[k] for k in keys])))
    D.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    grad_report('D bwd', list(D.named_parameters()), [q.grad for q in D.parameters()], gref)
    print('device error', tr.device_error())

    # ---------------- generator forward / backward ----------------
    taps = {}
    g = O._leaf(sdG)
    out_ref = O.generator_train(g, LR, f, 16, {}, taps)
    out = G(LR)
    for k in ['g_x0', 'g_x1', 'g_x8', 'g_x16', 'g_t', 'g_u0', 'g_u1', 'g_u2']:
        print(f'G fwd {k}: rel {rel(nhwc_to_nchw(tr.tensor(k)), taps[k]):.3e}')
    print(f'G out: rel {rel(out, out_ref):.3e}')
    gen = torch.Generator(device='cpu').manual_seed(9)
    dout = (torch.randn(out.shape, generator=gen) * 1e-3).to(dev)
    keys = O.param_keys(sdG)
    gref = dict(zip(keys, torch.autograd.grad((out_ref * dout).sum(), [g[k] for k in keys])))
    G.zero_grad()
    (out * dout).sum().backward()
    torch.cuda.synchronize()
    grad_report('G bwd', list(G.named_parameters()), [q.grad for q in G.parameters()], gref)
    print('device error', tr.device_error())

    # ---------------- perceptual loss ----------------
    fake = out_ref.detach().clone().requires_grad_(True)
    taps = {}
    f1 = O.vgg_features(sdV, O.vgg_transform(fake), taps)
    with torch.no_grad():
        f2 = O.vgg_features(sdV, O.vgg_transform(HR))
    l_ref = torch.nn.functional.mse_loss(f1, f2)
    (dfake_ref,) = torch.autograd.grad(l_ref, [fake])
    fake2 = out_ref.detach().clone().requires_grad_(True)
    l = V(fake2, HR)
    l.backward()
    torch.cuda.synchronize()
    print(f'VGG pre: rel {rel(nhwc_to_nchw(tr.tensor("v_pre"))[:, :3], O.vgg_transform(fake.detach())):.3e}')
    for i in (0, 1, 2, 3, 7, 11, 15):
        print(f'VGG fwd v_y{i}: rel {rel(nhwc_to_nchw(tr.tensor(f"v_y{i}")), taps[f"v_y{i}"]):.3e}')
    print(f'VGG loss {float(l):.6e} ref {float(l_ref):.6e}   dfake cos {cos(fake2.grad, dfake_ref):.5f} '
          f'norm ratio {float(fake2.grad.norm() / dfake_ref.norm()):.4f}')
    print('device error', tr.device_error())

    # ---------------- fused do_epoch against the oracle's ----------------
    torch.manual_seed(3)
    G2 = dsr_b200.Generator(f).to(dev).train()
    D2 = GT.Discriminator((H, W)).to(dev).train()
    step = GT.GanTrainStep(G2, D2, V, 1e-4, B, (lh, lw), dev)
    lD, lG = step.do_epoch(LR, HR)
    torch.cuda.synchronize()
    o = O.do_epoch(sdG, sdD, sdV, LR, HR, 1e-4, f)
    print(f'do_epoch loss_D {float(lD):.6f} ref {float(o["loss_D"]):.6f}   loss_G {float(lG):.6f} ref {float(o["loss_G"]):.6f}')
    grad_report('do_epoch gD', list(D2.named_parameters()), step.fd.grad_views, o['gD'])
    grad_report('do_epoch gG', list(G2.named_parameters()), step.fg.grad_views, o['gG'])
    for net, mod, sd in (('G', G2, sdG), ('D', D2, sdD)):
        worst = 0.0
        for k, v in mod.state_dict().items():
            if k.endswith('num_batches_tracked'):
                continue
            worst = max(worst, float((v - sd[k]).abs().max()))
        print(f'do_epoch post-step {net}: max |param - oracle| {worst:.3e} (lr 1e-4)')
    print('device error', tr.device_error())
    # ---------------- timing of the fused step ----------------
    for _ in range(3):
        step.do_epoch(LR, HR)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        step.do_epoch(LR, HR)
    e1.record()
    torch.cuda.synchronize()
    print(f'fused do_epoch: {e0.elapsed_time(e1) / n:.3f} ms per step (batch {B}, LR {lh}x{lw})')
    import time
    parts = {}
    fd, fg = step.fd, step.fg

    def timed(name, fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            r = fn()
        b.record()
        torch.cuda.synchronize()
        parts[name] = a.elapsed_time(b) / 5
        return r
    trn = step.tr
    timed('pack D', lambda: trn.pack(1, fd.flat, True))
    timed('pack G', lambda: trn.pack(0, fg.flat, True))
    timed('D fwd', lambda: trn.d_forward(0, fd.flat, fd.bflat, HR))
    fake = timed('G fwd', lambda: trn.g_forward(fg.flat, fg.bflat, LR, 1))
    timed('D bwd', lambda: trn.d_backward(0, fd.flat, fd.gflat, target=1.0))
    lossbuf = torch.zeros((), device=dev)
    dfake = timed('VGG loss+grad', lambda: trn.vgg_loss(fake, HR, lossbuf, False, True))
    timed('G bwd', lambda: trn.g_backward(fg.flat, dfake, fg.gflat))
    timed('Adam D', lambda: step._adam(fd.flat, fd.gflat, step.mD, step.vD))
    timed('Adam G', lambda: step._adam(fg.flat, fg.gflat, step.mG, step.vG))
    print('parts (ms):', {k: round(v, 3) for k, v in parts.items()})


if __name__ == '__main__':
    main()
