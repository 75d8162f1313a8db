"""The CPU oracle (oracle/dip_oracle.py) against fixtures produced by EXECUTING THE REFERENCE
(oracle/make_golden.py, run in the build container where /root/reference exists) and against the
operator known-answer tests of SURVEY.md section 4."""
import numpy as np
import pytest
import torch

from oracle import dip_oracle as O


def checksum(t):
    t = t.detach().double().flatten()
    return (float(t.sum()), float(t.abs().sum()),
            float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64)).sum() / max(1, t.numel())))


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize('f', [4, 8, 16])
def test_lanczos_table_matches_reference(golden, f):
    ref = golden('lanczos.pt')[f'kernel_f{f}'].numpy()
    k = O.lanczos_kernel_2d(f)
    assert k.shape == ref.shape == (4 * f, 4 * f)
    assert np.abs(k - ref).max() < 1e-16
    assert abs(k.sum() - 1) < 1e-14 and np.allclose(k, k.T)


def test_lanczos_kat_row_sums():
    # 1-D taps of get_kernel(4,'lanczos',0.5,17,support=2)
    want = [-0.001065, -0.009752, -0.020384, -0.014878, 0.024594, 0.098658, 0.183115, 0.239711]
    got = O.lanczos_kernel_2d(4).sum(axis=1)
    assert np.allclose(got[:8], want, atol=1e-6) and np.allclose(got[8:], want[::-1], atol=1e-6)


@pytest.mark.parametrize('f', [4, 8, 16])
def test_downsample_matches_reference(golden, f):
    g = golden('lanczos.pt')
    x = g['x'].clone().requires_grad_(True)
    y = O.downsample(x, f)
    assert y.shape == g[f'y_f{f}'].shape
    assert rel(y, g[f'y_f{f}']) < 1e-6
    (y * g[f'gy_f{f}']).sum().backward()
    assert rel(x.grad, g[f'gx_f{f}']) < 1e-6


def test_operator_kats():
    import torch.nn.functional as F
    a = torch.arange(5.).view(1, 1, 1, 5)
    assert F.pad(a, (1, 1, 0, 0), mode='reflect').flatten().tolist() == [1, 0, 1, 2, 3, 4, 3]
    b = torch.arange(4.).view(1, 1, 1, 4)
    assert F.pad(b, (2, 2, 0, 0), mode='replicate').flatten().tolist() == [0, 0, 0, 1, 2, 3, 3, 3]
    c = torch.arange(6.).view(1, 1, 1, 6).expand(1, 1, 2, 6)
    up = F.interpolate(c, scale_factor=2, mode='bilinear')[0, 0, 0]
    assert up[:4].tolist() == [0, 0.25, 0.75, 1.25] and up[-2:].tolist() == [4.75, 5.0]
    x = torch.randn(1, 3, 5, 7)
    y = O._bn(x, torch.ones(3), torch.zeros(3))
    m = x.mean(dim=(0, 2, 3), keepdim=True)
    v = ((x - m) ** 2).mean(dim=(0, 2, 3), keepdim=True)
    assert torch.allclose(y, (x - m) / torch.sqrt(v + 1e-5), atol=1e-6)


@pytest.mark.parametrize('seed', [0, 3])
def test_init_matches_reference(golden, seed):
    g = golden(f'init_seed{seed}.pt')
    threads = torch.get_num_threads()
    torch.set_num_threads(1)          # the fixture's float64 checksums were summed single-threaded
    try:
        torch.manual_seed(seed)
        sd = O.init_params()
        assert set(sd.keys()) == set(g['keys'])
        assert O.param_keys({k: sd[k] for k in g['keys']}) == g['param_order']
        for k, shape in g['shapes'].items():
            assert tuple(sd[k].shape) == tuple(shape), k
        for k, c in g['checksums'].items():
            assert checksum(sd[k]) == pytest.approx(c, rel=1e-12, abs=1e-12), k
    finally:
        torch.set_num_threads(threads)


@pytest.mark.parametrize('name', ['step_64x64.pt', 'step_64x96.pt', 'step_72x88.pt'])
def test_step_matches_reference(golden, name):
    fx = golden(name)
    torch.set_num_threads(1)          # the fixtures were generated single-threaded (fixes the reduction order)
    torch.manual_seed(fx['seed'])
    sd = O.init_params()
    loss, out, grads = O.step_loss_and_grads(sd, fx['z0'], fx['lr_img'], fx['factor'])
    assert rel(out, fx['out_hr']) < 1e-6      # bit-exact on the machine that generated the fixture
    assert float(loss) == pytest.approx(fx['losses'][0], rel=1e-5)
    assert rel(O.downsample(out, fx['factor']), fx['out_lr']) < 1e-5
    dead = set(O.dead_param_keys())
    # Besides the structurally dead parameters, a BatchNorm gamma whose output reaches the next train-mode
    # BatchNorm through positively homogeneous ops only (LeakyReLU, bilinear upsample) has a gradient that is pure
    # rounding noise while its beta is still 0 (first step): skip gradients 6 orders below the largest one.
    floor = 1e-6 * max(fx['grad_norms'].values())
    dead |= {k for k, n in fx['grad_norms'].items() if n < floor}
    for k, n in fx['grad_norms'].items():
        if k in dead:
            continue
        assert float(grads[k].double().norm()) == pytest.approx(n, rel=1e-4), k
    for k, g in fx['grad_full'].items():
        if k in dead:
            continue
        assert rel(grads[k], g) < 1e-4, k
    for k, g in fx['grad_slices'].items():
        assert rel(grads[k][:8, :8], g) < 1e-4, k
    # one Adam step (utils/DIP.py:33-38) reproduces the reference's post-step live parameters
    keys = O.param_keys(sd)
    adam = O.AdamState(keys, sd, fx['lr'])
    adam.step(sd, grads)
    for k, v in fx['post_adam_small'].items():
        if k in dead or k not in keys:
            continue
        # first Adam step is lr*sign(g): entries whose gradient is rounding noise may flip; compare the bulk
        close = (sd[k] - v).abs() < 1e-6
        assert close.float().mean() > 0.95, k


@pytest.mark.parametrize('name', ['step_zero_nearest_72x88.pt', 'step_zero_bilinear_64x64.pt',
                                  'step_reflection_nearest_64x96.pt'])
def test_step_matches_reference_other_pad_and_upsample_modes(golden, name):
    """pad='zero' (models/DIP/utils.py:96-102) and upsample_mode='nearest' (models/DIP/skip.py:77): the oracle follows
    the reference's fixtures (oracle/make_golden_modes.py), including the state_dict keys of the pad-less conv()."""
    fx = golden(name)
    torch.set_num_threads(1)
    torch.manual_seed(fx['seed'])
    sd = O.init_params(pad=fx['pad'])
    assert set(fx['keys']) == set(sd.keys())
    loss, out, grads = O.step_loss_and_grads(sd, fx['z0'], fx['lr_img'], fx['factor'], pad=fx['pad'],
                                             upsample_mode=fx['upsample_mode'])
    assert rel(out, fx['out_hr']) < 1e-6
    assert float(loss) == pytest.approx(fx['losses'][0], rel=1e-5)
    dead = set(O.dead_param_keys(pad=fx['pad']))
    floor = 1e-6 * max(fx['grad_norms'].values())
    dead |= {k for k, n in fx['grad_norms'].items() if n < floor}
    for k, n in fx['grad_norms'].items():
        if k not in dead:
            assert float(grads[k].double().norm()) == pytest.approx(n, rel=1e-4), k
    for k, g in fx['grad_full'].items():
        if k not in dead:
            assert rel(grads[k], g) < 1e-4, k
    for k, g in fx['grad_slices'].items():
        assert rel(grads[k][:8, :8], g) < 1e-4, k


def test_psnr_and_synthetic_pair():
    lr, hr = O.synthetic_pair(0, 64)
    assert hr.shape == (3, 64, 64) and lr.shape == (3, 16, 16)
    assert 0 <= float(hr.min()) and float(hr.max()) <= 1
    assert O.psnr(hr, hr + 0.1) == pytest.approx(20.0, abs=1e-3)


@pytest.mark.parametrize('name', ['step_256.pt', 'step_512.pt'])
def test_step_matches_reference_at_baseline_sizes(golden, name):
    """BASELINE configs[0] / configs[1] sizes: the oracle against one teacher-forced step of the unmodified reference
    (oracle/make_golden_large.py).  z0 is rebuilt from the seed and must hit the stored checksum."""
    from conftest import rebuild_z0
    from oracle.make_golden_large import probe
    fx = golden(name)
    torch.manual_seed(fx['seed'])
    sd = O.init_params()
    z0 = rebuild_z0(fx)
    assert checksum(z0) == pytest.approx(fx['z0_checksum'], rel=1e-9)
    loss, out, grads = O.step_loss_and_grads(sd, z0, fx['lr_img'], fx['factor'])
    assert rel(out, fx['out_hr']) < 1e-5          # multi-threaded reduction order differs from the fixture's: ~3e-7
    assert float(loss) == pytest.approx(fx['loss'], rel=1e-5)
    assert rel(O.downsample(out, fx['factor']), fx['out_lr']) < 1e-5
    dead = set(O.dead_param_keys())
    top = max(fx['grad_norms'].values())
    for k, n in fx['grad_norms'].items():
        if k in dead or n < 1e-6 * top:
            continue
        assert float(grads[k].double().norm()) == pytest.approx(n, rel=1e-3), k
        pr = float((grads[k].double() * probe(k, grads[k].shape)).sum())
        assert pr == pytest.approx(fx['grad_probe'][k], rel=2e-2, abs=2e-3 * n * grads[k].numel() ** 0.5), k
    for k, g in fx['grad_slices'].items():
        assert rel(grads[k][:8, :8], g) < 1e-3, k


def test_quantised_oracle_precision_class(golden):
    """What 16-bit operands cost on the freshly initialised (chaotic) network, measured on the CPU with the oracle's
    rounding hook at the points where the CUDA path stores 16-bit values (O.fp16_points): fp16 moves the output by
    ~4e-3 and the gradient to cosine ~0.99 of the fp32 reference, bf16 is an order of magnitude worse (why the CUDA
    path uses fp16 operands with a loss scale).  The GPU parity tests gate the CUDA path against THIS oracle."""
    fx = golden('step_64x64.pt')
    torch.manual_seed(fx['seed'])
    sd = O.init_params()
    _, out32, g32 = O.step_loss_and_grads(sd, fx['z0'], fx['lr_img'], fx['factor'])
    dead = set(O.dead_param_keys())
    live = [k for k in g32 if k not in dead and fx['grad_norms'][k] > 1e-6 * max(fx['grad_norms'].values())]

    def whole(g):
        return torch.cat([g[k].flatten() for k in live]).double()

    a = whole(g32)
    res = {}
    for tag, dt in (('fp16', torch.float16), ('bf16', torch.bfloat16)):
        loss, out, g = O.step_loss_and_grads(sd, fx['z0'], fx['lr_img'], fx['factor'], quant=O.fp16_points(dt))
        b = whole(g)
        res[tag] = (rel(out, out32), float(a @ b / (a.norm() * b.norm())), abs(float(loss) - fx['losses'][0]) / fx['losses'][0])
    assert res['fp16'][0] < 8e-3 and res['fp16'][1] > 0.985 and res['fp16'][2] < 1e-2     # measured 4.7e-3 / 0.992
    assert res['bf16'][0] > 3 * res['fp16'][0] and res['bf16'][1] < res['fp16'][1]        # measured 3.6e-2 / 0.45


def test_drop_in_downsampler_consumes_the_reference_rng_draws():
    """utils/downsampler.py:44 builds (and then overwrites) an nn.Conv2d: its initialisation advances the global CPU
    generator between get_net and get_noise (DIP.py:29,32).  The mirror draws the same numbers, so identical seeds
    give identical net_input / noise tensors."""
    import dsr_b200
    torch.manual_seed(5)
    dsr_b200.Downsampler(3, 4, 'lanczos2', phase=0.5, preserve_size=True)
    a = torch.rand(4)
    torch.manual_seed(5)
    torch.nn.Conv2d(3, 3, kernel_size=(16, 16), stride=4, padding=0)
    b = torch.rand(4)
    assert torch.equal(a, b)
