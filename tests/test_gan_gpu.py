"""GPU parity of the SRResNet generator inference path (dsr_b200.Generator -> dsr_gen_forward through the C ABI)
against the fixtures recorded from the unmodified reference (oracle/make_golden_gan.py) and against the CPU oracle."""
import ctypes as C

import pytest
import torch

from oracle import gan_oracle as go

pytestmark = pytest.mark.gpu
CASES = ['gan_f8_2x20x24.pt', 'gan_f8_1x17x23.pt', 'gan_f16_1x16x16.pt']
REL_L2_TOL = 1e-2          # north-star tolerance for network outputs under 16-bit operands / fp32 accumulation


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def build(fx):
    import dsr_b200
    torch.manual_seed(fx['seed'])
    g = dsr_b200.Generator(fx['factor'])
    sd = g.state_dict()
    go.perturb_trained_state(sd, fx['perturb_seed'])
    g.load_state_dict(sd)
    return g.cuda().eval()


@pytest.mark.parametrize('name', CASES)
def test_generator_matches_reference_fixture(golden, name):
    fx = golden(name)
    g = build(fx)
    y = g(fx['x'].cuda()).cpu()
    assert y.shape == fx['y'].shape
    assert torch.isfinite(y).all()
    assert rel_l2(y, fx['y']) <= REL_L2_TOL
    assert float((y - fx['y']).abs().max()) <= 5e-3          # outputs are tanh values in (-1, 1)


def test_generator_tap_by_tap_output_conv(golden, monkeypatch):
    """The 81-tap variant of the 9 x 9 output convolution (conv_halo2_kernel<2>, kept for A/B) gives the same image."""
    monkeypatch.setenv('DSR_GEN_CONV3_TAPS', '1')
    fx = golden('gan_f8_1x17x23.pt')
    y = build(fx)(fx['x'].cuda()).cpu()
    assert rel_l2(y, fx['y']) <= REL_L2_TOL


def test_generator_intermediates_match_oracle(golden):
    """Every stage (conv1, residual trunk, each PixelShuffle block) against the oracle's recorded intermediates."""
    from dsr_b200._lib import lib, check
    fx = golden('gan_f8_2x20x24.pt')
    g = build(fx)
    x = fx['x']
    g(x.cuda())
    torch.cuda.synchronize()
    rec = {}
    go.generator_forward({k: v.cpu() for k, v in g.state_dict().items()}, x, fx['factor'], record=rec)
    plan = next(iter(g._plans.values()))
    B = x.shape[0]

    def fetch(name):
        ptr, rows, img_rows, H, W = C.c_void_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib.dsr_gen_debug_tensor(plan.handle, name.encode(), C.byref(ptr), C.byref(rows), C.byref(img_rows),
                                       C.byref(H), C.byref(W)))
        t = torch.empty((rows.value, W.value, 64), dtype=torch.float16, device='cuda')
        check(lib.dsr_debug_copy(t.data_ptr(), ptr, t.numel() * 2, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        t = t.cpu().float()
        imgs = [t[b * img_rows.value: b * img_rows.value + H.value] for b in range(B)]
        gaps = [t[b * img_rows.value + H.value: (b + 1) * img_rows.value] for b in range(B - 1)]
        for gp in gaps:
            assert float(gp.abs().max()) == 0.0, f'{name}: rows between images must stay zero'
        return torch.stack(imgs).permute(0, 3, 1, 2)

    assert rel_l2(fetch('x0'), rec['x0']) <= 2e-3
    assert rel_l2(fetch('t1'), rec['block15_t']) <= REL_L2_TOL
    nsh = go.SHUFFLES[fx['factor']]
    for i in range(nsh):
        assert rel_l2(fetch(f's{i}'), rec[f's{i}']) <= REL_L2_TOL, f's{i}'


def test_generator_protocol_errors():
    import dsr_b200
    g = dsr_b200.Generator(8).cuda()
    y = g(torch.rand(1, 3, 16, 16, device='cuda'))          # training mode (train_GAN.py): batch-statistics BatchNorm
    assert y.shape == (1, 3, 128, 128) and y.requires_grad and bool(torch.isfinite(y).all())
    with pytest.raises(NotImplementedError):
        g(torch.rand(1, 3, 16, 16, device='cuda').requires_grad_(True))   # gradient w.r.t. the LR input is not built
    with pytest.raises(RuntimeError):
        g.eval()(torch.rand(1, 3, 16, 16))                   # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        dsr_b200.Generator(4)                                # the reference cannot build x4 either


def test_generator_weights_reload_and_chunking(golden):
    """load_state_dict after a forward is picked up, and a batch larger than max_chunk equals per-image calls."""
    fx = golden('gan_f8_1x17x23.pt')
    g = build(fx)
    x = torch.rand(5, 3, 17, 23, generator=torch.Generator().manual_seed(3)).cuda()
    g.max_chunk = 2
    y = g(x)
    g.max_chunk = 8
    y1 = g(x)
    assert rel_l2(y.cpu(), y1.cpu()) <= 1e-6
    sd = {k: v.clone() for k, v in g.state_dict().items()}
    sd['conv3.bias'] += 0.25
    g.load_state_dict(sd)
    y2 = g(x)
    assert float((y2 - y1).abs().mean()) > 1e-2


def test_generator_config4_patch_size_matches_oracle():
    """BASELINE configs[3] geometry (96 x 96 LR patches, x8 -> 768 x 768), three images, against the CPU oracle."""
    import dsr_b200
    torch.manual_seed(21)
    g = dsr_b200.Generator(8)
    sd = g.state_dict()
    go.perturb_trained_state(sd, 4)
    g.load_state_dict(sd)
    x = torch.rand(3, 3, 96, 96, generator=torch.Generator().manual_seed(8))
    want = go.generator_forward({k: v.clone() for k, v in g.state_dict().items()}, x, 8)
    got = g.cuda().eval()(x.cuda()).cpu()
    assert got.shape == (3, 3, 768, 768)
    assert rel_l2(got, want) <= REL_L2_TOL
    # images of a batch are independent: permuting the batch permutes the outputs (tall-grid gaps hold)
    perm = torch.tensor([2, 0, 1])
    got_p = g(x[perm].cuda()).cpu()
    assert rel_l2(got_p, got[perm]) <= 1e-6


def test_generator_plan_cache_is_bounded():
    """eval_GAN.py feeds images of many sizes: plans (and their workspaces) beyond max_plans are dropped, LRU first."""
    import dsr_b200
    g = dsr_b200.Generator(8).cuda().eval()
    g.max_plans = 2
    outs = {}
    for hw in [(16, 16), (16, 24), (24, 16), (16, 16)]:
        x = torch.rand(1, 3, *hw, generator=torch.Generator().manual_seed(hw[0] * 100 + hw[1])).cuda()
        y = g(x)
        assert y.shape == (1, 3, 8 * hw[0], 8 * hw[1]) and torch.isfinite(y).all()
        if hw in outs:
            assert rel_l2(y.cpu(), outs[hw]) <= 1e-6       # a re-created plan gives the same image
        outs[hw] = y.cpu()
        assert len(g._plans) <= 2
