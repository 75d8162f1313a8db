"""SRGAN training step on the GPU (BASELINE configs[4]) against the oracle, the reference
fixtures and the reference's own training loop.  bf16 operands / fp32 accumulation: the gates are the measured bf16
class, stated per assertion."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cos(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


def whole(named, grads, ref, skip=()):
    a = torch.cat([g.flatten().double() for (k, _), g in zip(named, grads) if k not in skip])
    b = torch.cat([ref[k].flatten().double().to(a.device) for k, _ in named if k not in skip])
    return cos(a, b), float(a.norm() / b.norm())


@pytest.fixture(scope='module')
def env():
    import dsr_b200
    from dsr_b200 import gan_train as GT
    from oracle import gan_train_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return dsr_b200, GT, O, torch.device('cuda:0')


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def test_discriminator_forward_backward_full_size(env):
    """Discriminator (discriminator.py:57-74) on a batch of 8 192x192 patches: activations, probabilities and the
    gradient of BCE(D(x), 1) against the oracle in fp32 on the same device."""
    dsr_b200, GT, O, dev = env
    torch.manual_seed(3)
    D = GT.Discriminator((192, 192)).to(dev).train()
    _, HR = O.synthetic_batch(5, 8, (24, 24), 8)
    HR = HR.to(dev)
    sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    d = O._leaf(sd)
    taps, ns = {}, {}
    p_ref = O.discriminator_train(d, HR, ns, taps)
    p = D(HR)
    tr = D._trainer_for(HR)[0]
    # precision-class yardstick: the ORACLE with bf16 rounding at the same storage points lands just as far from fp32
    # (rounding noise of ~20 storage points, decorrelated between two realisations by the first flipped rounding), so the
    # CUDA path may not be further from fp32 than 1.3 x that -- a kernel fault cannot hide inside "it is only bf16"
    tq = {}
    p_q = O.discriminator_train(sd, HR, {}, tq, q=O.bf16_points())
    e32, eq = rel(nchw(tr.tensor('d_h7')), taps['d_h7']), rel(tq['d_h7'], taps['d_h7'])
    print(f'D: d_h7 rel {e32:.2e} vs the fp32 oracle; the bf16-point oracle itself is {eq:.2e} from fp32 '
          f'(CUDA vs bf16-point oracle {rel(nchw(tr.tensor("d_h7")), tq["d_h7"]):.2e})')
    assert e32 < 1.3 * eq and float((p - p_ref).abs().max()) < 1.5 * float((p_q - p_ref).abs().max()) + 1e-3
    assert rel(nchw(tr.tensor('d_h0')), taps['d_h0']) < 6e-3            # one bf16 rounding
    assert rel(nchw(tr.tensor('d_h4')), taps['d_h4']) < 2.5e-2
    assert rel(nchw(tr.tensor('d_h7')), taps['d_h7']) < 3.5e-2          # eight bf16 layers deep
    assert float((p - p_ref).abs().max()) < 5e-3
    keys = O.param_keys(sd)
    gref = dict(zip(keys, torch.autograd.grad(O.bce(p_ref, 1.0), [d[k] for k in keys])))
    D.zero_grad()
    O.bce(p, 1.0).backward()
    named = list(D.named_parameters())
    c, r = whole(named, [q.grad for _, q in named], gref)
    print(f'D: prob max err {float((p - p_ref).abs().max()):.2e}  whole-gradient cosine {c:.4f} norm ratio {r:.4f}')
    assert c > 0.99 and abs(r - 1) < 2e-2
    for k, q in named:                                                   # every live tensor
        if float(gref[k].norm()) > 1e-6 * float(gref['dense1.weight'].norm()) and not k.endswith('conv1.bias'):
            assert cos(q.grad, gref[k]) > 0.95, k
    # running statistics of the train-mode pass (momentum 0.1, unbiased variance, conv bias included in the mean)
    for k, v in ns.items():
        assert torch.allclose(D.state_dict()[k], v, rtol=2e-2, atol=2e-3), k
    assert tr.device_error() == 0


def test_generator_train_forward_backward_full_size(env):
    """Generator in train mode (generator.py:68-81) on 8 24x24 LR patches and its backward pass."""
    dsr_b200, GT, O, dev = env
    torch.manual_seed(3)
    G = dsr_b200.Generator(8).to(dev).train()
    LR, _ = O.synthetic_batch(5, 8, (24, 24), 8)
    LR = LR.to(dev)
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    g = O._leaf(sd)
    taps, ns = {}, {}
    out_ref = O.generator_train(g, LR, 8, 16, ns, taps)
    out = G(LR)
    tr = G._train_state(LR)[0]
    assert rel(nchw(tr.tensor('g_x0')), taps['g_x0']) < 6e-3
    assert rel(nchw(tr.tensor('g_x16')), taps['g_x16']) < 2.5e-2
    assert rel(nchw(tr.tensor('g_u2')), taps['g_u2']) < 2.5e-2
    assert rel(out, out_ref) < 2.5e-2                                     # 38 bf16 layers; 1.5e-2 measured
    with torch.no_grad():
        out_q = O.generator_train(sd, LR, 8, 16, q=O.bf16_points())
    print(f'G: out rel {rel(out, out_ref):.2e} vs the fp32 oracle; the bf16-point oracle itself is {rel(out_q, out_ref):.2e} '
          f'from fp32 (CUDA vs bf16-point oracle {rel(out, out_q):.2e})')
    assert rel(out, out_ref) < 1.3 * rel(out_q, out_ref)                  # the precision-class yardstick (see the D test)
    dout = (torch.randn(out.shape, generator=torch.Generator().manual_seed(9)) * 1e-3).to(dev)
    keys = O.param_keys(sd)
    gref = dict(zip(keys, torch.autograd.grad((out_ref * dout).sum(), [g[k] for k in keys])))
    G.zero_grad()
    (out * dout).sum().backward()
    named = list(G.named_parameters())
    c, r = whole(named, [q.grad for _, q in named], gref)
    print(f'G: out rel {rel(out, out_ref):.2e}  whole-gradient cosine {c:.4f} norm ratio {r:.4f}')
    assert c > 0.985 and abs(r - 1) < 2e-2
    for k in ('conv1.weight', 'conv3.weight', 'conv3.bias', 'pixel_shuffle_blocks.2.conv1.weight',
              'pixel_shuffle_blocks.0.conv1.bias', 'residual_blocks.0.conv1.weight', 'residual_blocks.15.bn2.weight'):
        assert cos(dict(named)[k].grad, gref[k]) > 0.95, k
    for k, v in ns.items():
        assert torch.allclose(G.state_dict()[k], v, rtol=2e-2, atol=2e-3), k
    assert tr.device_error() == 0


def test_perceptual_loss_and_gradient(env):
    """Vgg19Loss (utils/GAN.py:62-88): transform, 16 convolutions, 4 pools, feature MSE, gradient w.r.t. the image."""
    dsr_b200, GT, O, dev = env
    torch.manual_seed(11)
    V = GT.Vgg19Loss(pretrained=False).to(dev)
    sdV = {k: v.detach().clone() for k, v in V.state_dict().items()}
    g = torch.Generator().manual_seed(12)
    real = torch.rand(4, 3, 64, 64, generator=g).to(dev)
    fake = (real + 0.2 * torch.randn(real.shape, generator=g).to(dev)).clamp(-1, 1)
    f_ref = fake.clone().requires_grad_(True)
    taps = {}
    l_ref = torch.nn.functional.mse_loss(O.vgg_features(sdV, O.vgg_transform(f_ref), taps),
                                         O.vgg_features(sdV, O.vgg_transform(real)))
    (d_ref,) = torch.autograd.grad(l_ref, [f_ref])
    f = fake.clone().requires_grad_(True)
    loss = V(f, real)
    loss.backward()
    tr = V._trainer_for(real)
    assert rel(nchw(tr.tensor('v_pre'))[:, :3], O.vgg_transform(fake)) < 4e-3
    assert rel(nchw(tr.tensor('v_y0')), taps['v_y0']) < 6e-3
    assert rel(nchw(tr.tensor('v_y15')), taps['v_y15']) < 2e-2
    print(f'VGG: loss {float(loss):.5e} vs {float(l_ref):.5e}  d/dfake cosine {cos(f.grad, d_ref):.4f} '
          f'norm ratio {float(f.grad.norm() / d_ref.norm()):.4f}')
    assert abs(float(loss) - float(l_ref)) < 2e-2 * float(l_ref)
    assert cos(f.grad, d_ref) > 0.93 and abs(float(f.grad.norm() / d_ref.norm()) - 1) < 5e-2
    assert tr.device_error() == 0


@pytest.mark.parametrize('case', ['gan_train_2x8x8.pt', 'gan_train_3x16x8.pt'])
def test_fused_step_against_reference_fixture(env, golden, case):
    """GanTrainStep.do_epoch from the fixture's seeds against what the UNMODIFIED reference's do_epoch produced
    (losses, generator gradient) and against the oracle's full gradients."""
    dsr_b200, GT, O, dev = env
    fx = golden(case)
    h, w = fx['lr_hw']
    torch.manual_seed(fx['seed'])
    G = dsr_b200.Generator(fx['factor'])
    D = GT.Discriminator((h * fx['factor'], w * fx['factor']))
    torch.manual_seed(fx['vgg_seed'])
    V = GT.Vgg19Loss(pretrained=False)
    for k, c in fx['init']['D'].items():                  # same-seed initial state incl. fc_input_shape's statistics
        t = D.state_dict()[k].double().flatten()
        assert (float(t.sum()), float(t.abs().sum())) == pytest.approx(c, rel=1e-5, abs=1e-7), k   # CPU conv of fc_input_shape
    sdG = {k: v.detach().clone().to(dev) for k, v in G.state_dict().items()}
    sdD = {k: v.detach().clone().to(dev) for k, v in D.state_dict().items()}
    sdV = {k: v.detach().clone().to(dev) for k, v in V.state_dict().items()}
    step = GT.GanTrainStep(G.train(), D.train(), V.to(dev), fx['lr'], fx['batch'], (h, w), dev)
    lD, lG = step.do_epoch(fx['LR'], fx['HR'])
    torch.cuda.synchronize()
    print(f'{case}: loss_D {float(lD):.5f} (reference {fx["loss_D"]:.5f})  loss_G {float(lG):.5f} (reference {fx["loss_G"]:.5f})')
    assert abs(float(lD) - fx['loss_D']) < 1e-2 * fx['loss_D']
    assert abs(float(lG) - fx['loss_G']) < 2e-2 * fx['loss_G']
    o = O.do_epoch(sdG, sdD, sdV, fx['LR'].to(dev), fx['HR'].to(dev), fx['lr'], fx['factor'])
    namedG, namedD = list(G.named_parameters()), list(D.named_parameters())
    dead = [k for k, _ in namedG + namedD if k.endswith(('conv1.bias', 'conv2.bias')) and ('blocks' in k or k == 'conv2.bias')]
    cG, rG = whole(namedG, step.fg.grad_views, o['gG'], dead)
    cD, rD = whole(namedD, step.fd.grad_views, o['gD'], dead)
    print(f'   whole-gradient cosine G {cG:.4f} (norm ratio {rG:.4f})  D {cD:.4f} ({rD:.4f})')
    assert cG > 0.99 and cD > 0.97 and abs(rG - 1) < 3e-2 and abs(rD - 1) < 3e-2
    # generator gradient checksums of the reference itself (sum of |g| per tensor), the large tensors
    for k in ('conv1.weight', 'conv2.weight', 'conv3.weight', 'pixel_shuffle_blocks.1.conv1.weight'):
        got = float(dict(zip([n for n, _ in namedG], step.fg.grad_views))[k].abs().sum())
        assert got == pytest.approx(fx['grad_G'][k][0][1], rel=5e-2), k
    assert step.tr.device_error() == 0


def test_workspace_invariants_after_training_steps(env):
    """The tall-grid layout rests on one invariant: the gap rows between the images of a batch (the convolutions' zero
    padding) and the padded channels of the 3-channel tensors are never written.  After three fused steps every
    inspectable tensor still has zero gaps, the canary behind the workspace is intact and no kernel reported a stalled
    barrier.  (compute-sanitizer is not available on this pool.)"""
    dsr_b200, GT, O, dev = env
    torch.manual_seed(2)
    G, D = dsr_b200.Generator(8).train(), GT.Discriminator((64, 96)).train()
    V = GT.Vgg19Loss(pretrained=False).to(dev)
    step = GT.GanTrainStep(G, D, V, 1e-4, 3, (8, 12), dev)                 # non-square, widths 12 .. 96
    LR, HR = O.synthetic_batch(3, 3, (8, 12), 8)
    for _ in range(3):
        lD, lG = step.do_epoch(LR, HR)
    torch.cuda.synchronize()
    tr = step.tr
    assert bool(torch.isfinite(lD)) and bool(torch.isfinite(lG))
    names = ['g_z1', 'g_x0', 'g_x7', 'g_x16', 'g_t', 'g_u0', 'g_u1', 'g_u2', 'g_z', 'd_h0', 'd_h1', 'd_h4', 'd_h7', 'v_pre',
             'v_y0', 'v_y1', 'v_y5', 'v_y15']
    for n in names:
        full, valid = tr.tensor(n, with_gap=True), tr.tensor(n)
        assert float(full[:, valid.shape[1]:].abs().max()) == 0.0, f'{n}: gap rows written'
        assert float(valid.abs().max()) > 0.0, n
    assert float(tr.tensor('v_pre')[..., 3:].abs().max()) == 0.0 and float(tr.tensor('g_z')[..., 3:].abs().max()) == 0.0
    assert tr.guard_intact() and tr.device_error() == 0


def test_three_stream_step_equals_the_single_stream_order(env, monkeypatch):
    """GanTrainStep runs the discriminator chain, the generator chain and the real-batch VGG pass on three streams
    (DESIGN.md 10 Streams); DSR_GAN_ONE_STREAM=1 issues the same calls on one.  Same seeds, same batch: the losses and
    the gradients of both networks after the first step, and the losses of the second step (which see the first step's
    Adam updates and BatchNorm statistics), agree to the noise of the fp32 atomics' order -- a missing event between
    the streams (a race on `fake`, on the packed weights, on the BatchNorm scratch) would not."""
    dsr_b200, GT, O, dev = env
    LR, HR = O.synthetic_batch(41, 4, (16, 16), 8)
    runs = []
    for one_stream in (True, True, False):              # the single-stream order twice: the run-to-run yardstick
        if one_stream:
            monkeypatch.setenv('DSR_GAN_ONE_STREAM', '1')
        else:
            monkeypatch.delenv('DSR_GAN_ONE_STREAM', raising=False)
        torch.manual_seed(23)
        G, D = dsr_b200.Generator(8).train(), GT.Discriminator((128, 128)).train()
        torch.manual_seed(1023)
        V = GT.Vgg19Loss(pretrained=False).to(dev)
        step = GT.GanTrainStep(G, D, V, 1e-4, 4, (16, 16), dev)
        assert (step._side is None) == one_stream
        lD, lG = step.do_epoch(LR, HR)
        torch.cuda.synchronize()
        first = (float(lD), float(lG), step.fg.gflat.clone(), step.fd.gflat.clone())
        for _ in range(2):
            lD, lG = step.do_epoch(LR, HR)
        torch.cuda.synchronize()
        runs.append(first + (float(lD), float(lG)))
        assert step.tr.device_error() == 0
    a0, a, b = runs

    def show(tag, x, y):
        print(f'{tag}: loss_D {x[0]:.6f} / {y[0]:.6f}  loss_G {x[1]:.6f} / {y[1]:.6f}  gradient cosine G {cos(x[2], y[2]):.7f} '
              f'D {cos(x[3], y[3]):.7f}; third step loss_D {x[4]:.6f} / {y[4]:.6f}  loss_G {x[5]:.6f} / {y[5]:.6f}')
    show('one stream, run 1 / run 2', a0, a)
    show('one stream / three streams', a, b)
    # loss_D and both gradients of the first step come before any update: equal up to the order of the fp32 atomics.
    # loss_G of the first step already holds BCE(D'(fake)) after Adam(D), which turns that noise into +-lr steps.
    assert a[0] == pytest.approx(b[0], rel=1e-5)
    assert cos(a[2], b[2]) > 0.99999 and cos(a[3], b[3]) > 0.9999
    assert a[1] == pytest.approx(b[1], rel=max(2e-3, 5 * abs(a0[1] - a[1]) / abs(a[1])))
    assert a[4] == pytest.approx(b[4], rel=max(0.15, 5 * abs(a0[4] - a[4]) / abs(a[4])))     # measured run to run: 2 %
    assert a[5] == pytest.approx(b[5], rel=max(0.05, 5 * abs(a0[5] - a[5]) / abs(a[5])))     # measured run to run: 0.3 %


def test_factor16_step_with_12x12_patches(env):
    """train_GAN.py --downsample: factor 16, HR patch 192 -> LR patch 12 x 12 (train_GAN.py:240-270); four PixelShuffle
    blocks and level widths that are no multiple of the 8-pixel tiles.  One fused step against the oracle."""
    dsr_b200, GT, O, dev = env
    torch.manual_seed(17)
    G = dsr_b200.Generator(16).train()
    D = GT.Discriminator((192, 192)).train()
    torch.manual_seed(1017)
    V = GT.Vgg19Loss(pretrained=False)
    sdG, sdD, sdV = ({k: v.detach().clone().to(dev) for k, v in m.state_dict().items()} for m in (G, D, V))
    LR, HR = O.synthetic_batch(18, 2, (12, 12), 16)
    step = GT.GanTrainStep(G, D, V.to(dev), 1e-4, 2, (12, 12), dev)
    lD, lG = step.do_epoch(LR, HR)
    torch.cuda.synchronize()
    o = O.do_epoch(sdG, sdD, sdV, LR.to(dev), HR.to(dev), 1e-4, 16)
    namedG, namedD = list(G.named_parameters()), list(D.named_parameters())
    dead = [k for k, _ in namedG + namedD if k.endswith(('conv1.bias', 'conv2.bias')) and ('blocks' in k or k == 'conv2.bias')]
    cG, rG = whole(namedG, step.fg.grad_views, o['gG'], dead)
    cD, rD = whole(namedD, step.fd.grad_views, o['gD'], dead)
    print(f'x16: loss_D {float(lD):.5f} vs {float(o["loss_D"]):.5f}  loss_G {float(lG):.5f} vs {float(o["loss_G"]):.5f}  '
          f'gradient cosine G {cG:.4f} D {cD:.4f}')
    assert abs(float(lD) - float(o['loss_D'])) < 1e-2 * float(o['loss_D'])
    assert abs(float(lG) - float(o['loss_G'])) < 2e-2 * float(o['loss_G'])
    assert cG > 0.99 and cD > 0.97 and abs(rG - 1) < 3e-2 and abs(rD - 1) < 3e-2
    assert step.tr.device_error() == 0


@pytest.mark.skipif(not os.path.isdir(REF), reason='baseline/_ref not staged (no /root/reference at build time)')
def test_reference_training_loop_over_the_drop_in_modules():
    """The reference's own GAN_ISR_train / do_epoch (unmodified, baseline/_ref/train_GAN.py) with loss.backward() and
    torch.optim.Adam over the drop-in Generator / Discriminator / PerceptualLoss, next to the same loop over the
    reference's own modules in stock fp32 CUDA eager: same seeds, three steps."""
    def run(impl):
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'run_reference_gan.py'), '--impl', impl,
                              '--device', 'cuda', '--batch', '8', '--lr-size', '24', '--epochs', '3', '--lr', '1e-5'],
                             capture_output=True, text=True, check=True).stdout
        return json.loads(out.strip().splitlines()[-1])
    ours, ref = run('ours'), run('reference')
    print(f"reference do_epoch x3: over dsr_b200 {ours['steps_per_s']:.2f} steps/s incl. set-up, losses "
          f"{ours['loss_D']:.4f} / {ours['loss_G']:.4f};  over its own modules (fp32 eager) {ref['steps_per_s']:.2f} steps/s, "
          f"{ref['loss_D']:.4f} / {ref['loss_G']:.4f}")
    assert ours['modules'].startswith('deep-super-resolution_b200/') and ref['modules'].startswith('baseline/_ref/')
    assert abs(ours['loss_D'] - ref['loss_D']) < 0.1 * ref['loss_D']        # after two Adam steps of each network
    assert abs(ours['loss_G'] - ref['loss_G']) < 3e-2 * ref['loss_G']
    assert abs(ours['psnr0'] - ref['psnr0']) < 0.1
    assert ours['bn_mean_abs'] == pytest.approx(ref['bn_mean_abs'], rel=5e-2)
