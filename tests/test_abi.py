"""Host-side checks that need no GPU: the C-ABI library loads and exports exactly the symbols that
include/dsr_b200.h declares, the host-only entry points (Lanczos table, plan layout) agree with the
reference fixtures, and the Python mirror refuses unsupported configurations / CPU tensors."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'dsr_b200.h')


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dsr_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from dsr_b200 import _lib
    declared = header_symbols()
    assert len(declared) >= 28
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r' T (dsr_[a-z0-9_]+)', out)))
    assert exported == declared
    assert sorted(_lib.SIGNATURES) == declared          # the ctypes table binds all of them
    assert _lib.lib.dsr_abi_version() == 1
    assert _lib.lib.dsr_error_string(-5).decode() == 'unsupported network configuration'


def test_library_is_sm100a_tcgen05():
    """The shipped cubin is sm_100a and holds the Blackwell tensor/TMA instructions (SASS mnemonics)."""
    from dsr_b200 import _lib
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in sass
    assert 'UTCHMMA' in sass and 'UTMALDG' in sass and 'LDTM' in sass


@pytest.mark.parametrize('f', [4, 8, 16])
def test_lanczos_table(golden, f):
    import dsr_b200
    k = dsr_b200.get_kernel(f, 'lanczos', 0.5, 4 * f + 1, support=2)
    ref = golden('lanczos.pt')[f'kernel_f{f}'].numpy()
    assert k.shape == ref.shape and np.abs(k - ref).max() < 1e-15
    ds = dsr_b200.Downsampler(3, f, 'lanczos2', phase=0.5, preserve_size=True)
    assert np.array_equal(ds.kernel, k) and ds.out_size(64, 96) == (64 // f, 96 // f)


def test_plan_layout_matches_reference_state_dict(golden):
    from dsr_b200._lib import lib, check
    from dsr_b200.net import _read_layout
    g = golden('init_seed0.pt')
    h = C.c_void_p()
    check(lib.dsr_plan_create(C.byref(h), 64, 96, 32, 5, 3))
    lay = _read_layout(h)
    assert [n for n, _, _ in lay['params']] == g['param_order']
    assert lay['nparam'] == 2217831 and len(lay['params']) == 112 and len(lay['bns']) == 30
    off = 0
    for name, o, shape in lay['params']:
        assert o == off and tuple(shape) == tuple(g['shapes'][name]), name
        off += int(np.prod(shape))
    bn_names = [k[:-len('.running_mean')] for k in g['keys'] if k.endswith('.running_mean')]
    assert [n for n, _, _ in lay['bns']] == bn_names
    assert lib.dsr_plan_workspace_bytes(h) > 0
    lib.dsr_plan_destroy(h)
    check(lib.dsr_plan_create(C.byref(h), 127, 84, 32, 5, 3))      # odd level sizes are supported (Concat crop)
    lib.dsr_plan_destroy(h)
    for bad in ((20, 64, 32, 5, 3), (64, 64, 24, 5, 3), (64, 64, 32, 7, 3), (64, 64, 32, 5, 1)):
        assert lib.dsr_plan_create(C.byref(h), *bad) == -5


@pytest.mark.parametrize('seed', [0, 3])
def test_get_net_same_seed_init_and_state_dict(golden, seed):
    import dsr_b200
    g = golden(f'init_seed{seed}.pt')
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        torch.manual_seed(seed)
        net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                               upsample_mode='bilinear')
        sd = net.state_dict()
        assert list(sd.keys()) == g['keys']
        assert [n for n, _ in net.named_parameters()] == g['param_order']
        for k, c in g['checksums'].items():
            t = sd[k].detach().double().flatten()
            got = (float(t.sum()), float(t.abs().sum()),
                   float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64)).sum() / max(1, t.numel())))
            assert got == pytest.approx(c, rel=1e-12, abs=1e-12), k
        assert net.training
        # parity injection path: loading a reference-shaped state_dict works strictly
        net.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    finally:
        torch.set_num_threads(threads)


def test_unsupported_configurations_raise():
    import dsr_b200
    with pytest.raises(NotImplementedError):
        dsr_b200.get_net(32, 'skip', 'replicate', 'bilinear')
    with pytest.raises(NotImplementedError):
        dsr_b200.get_net(32, 'skip', 'reflection', 'bicubic')
    with pytest.raises(NotImplementedError):
        dsr_b200.get_net(32, 'skip', 'reflection', 'bilinear', downsample_mode='avg')
    with pytest.raises(NotImplementedError):
        dsr_b200.get_net(32, 'skip', 'reflection', 'bilinear', skip_n33d=64)
    with pytest.raises(NotImplementedError):
        dsr_b200.Downsampler(3, 4, 'gauss12', phase=0.5, preserve_size=True)
    with pytest.raises(NotImplementedError):
        dsr_b200.Downsampler(3, 4, 'lanczos2', phase=0, preserve_size=True)
    with pytest.raises(NotImplementedError):
        dsr_b200.optimize('LBFGS', [], lambda: None, 0.01, 1)
    with pytest.raises(NotImplementedError):
        dsr_b200.get_noise(2, 'meshgrid', (8, 8))


def test_pad_zero_and_nearest_keep_the_reference_state_dict_keys():
    """get_net(pad='zero') drops conv()'s padder module, so the Conv2d becomes child '0' of its Sequential
    (models/DIP/utils.py:96-105); the drop-in follows, and same-seed initial values are unchanged."""
    import dsr_b200
    torch.manual_seed(4)
    a = dsr_b200.get_net(32, 'skip', 'reflection', 'bilinear', num_scales=2)
    torch.manual_seed(4)
    b = dsr_b200.get_net(32, 'skip', 'zero', 'nearest', num_scales=2)
    ka, kb = list(a.state_dict().keys()), list(b.state_dict().keys())
    assert len(ka) == len(kb) and '1.0.1.1.weight' in ka and '1.0.1.0.weight' in kb and '9.0.bias' in kb
    for x, y in zip(ka, kb):
        assert torch.equal(a.state_dict()[x], b.state_dict()[y])


def test_no_cpu_fallback():
    import dsr_b200
    net = dsr_b200.get_net(32, 'skip', 'reflection', 'bilinear')
    with pytest.raises(RuntimeError, match='CUDA'):
        net(torch.zeros(1, 32, 64, 64))
    ds = dsr_b200.Downsampler(3, 4, 'lanczos2', phase=0.5, preserve_size=True)
    with pytest.raises(RuntimeError, match='CUDA'):
        ds(torch.zeros(1, 3, 64, 64))
    with pytest.raises(NotImplementedError):
        dsr_b200.optimize('adam', [torch.nn.Parameter(torch.zeros(3))], lambda: None, 0.01, 1)


def test_get_noise_matches_reference_recipe(monkeypatch):
    import dsr_b200
    monkeypatch.setenv('DSR_NOISE_DEVICE', 'cpu')
    torch.manual_seed(4)
    z = dsr_b200.get_noise(32, 'noise', (16, 24))
    torch.manual_seed(4)
    want = torch.zeros(1, 32, 16, 24).uniform_() * 0.1          # utils/DIP.py:92-96
    assert torch.equal(z, want) and z.device.type == 'cpu'


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'deep-super-resolution_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'(from|import)\s+oracle|dip_oracle|oracle/', src), (dirpath, f)


def test_drop_in_module_names_resolve_to_this_package():
    import importlib
    import dsr_b200
    assert importlib.import_module('models.DIP').get_net is dsr_b200.get_net
    assert importlib.import_module('utils.downsampler').Downsampler is dsr_b200.Downsampler
    m = importlib.import_module('utils.DIP')
    assert m.optimize is dsr_b200.optimize and m.get_noise is dsr_b200.get_noise and m.get_params is dsr_b200.get_params
    common = importlib.import_module('utils.common')
    assert os.path.realpath(common.__file__).startswith(os.path.realpath(os.path.join(ROOT, 'deep-super-resolution_b200')))
    assert m.save_image is common.save_image and m.np_to_pil is common.np_to_pil     # utils/DIP.py:3 re-exports utils.common


# ---- SRResNet generator (dsr_gen_*): host-side checks, no GPU work ------------------------------------------
@pytest.mark.parametrize('name', ['gan_f8_1x17x23.pt', 'gan_f16_1x16x16.pt'])
def test_generator_plan_layout_and_same_seed_init(golden, name):
    """The C plan's flat state layout lists the reference module's float state_dict entries in order, and
    dsr_b200.Generator reproduces the reference's same-seed initialisation (checksums recorded from the reference)."""
    import ctypes as C
    import torch
    import dsr_b200
    from dsr_b200._lib import lib, check
    from oracle import gan_oracle as go
    fx = golden(name)
    torch.manual_seed(fx['seed'])
    g = dsr_b200.Generator(fx['factor'])
    sd = g.state_dict()
    assert list(sd.keys()) == fx['keys']
    go.perturb_trained_state(sd, fx['perturb_seed'])
    for k, v in sd.items():
        t = v.detach().double().flatten()
        assert (float(t.sum()), float(t.abs().sum())) == pytest.approx(fx['checksums'][k], rel=1e-12, abs=1e-12), k
    plan = C.c_void_p()
    check(lib.dsr_gen_plan_create(C.byref(plan), fx['factor'], 16, 2, 24, 32))
    try:
        names, total = [], 0
        buf = C.create_string_buffer(128)
        off, n = C.c_longlong(), C.c_longlong()
        for i in range(lib.dsr_gen_num_tensors(plan)):
            check(lib.dsr_gen_tensor_info(plan, i, buf, 128, C.byref(off), C.byref(n)))
            assert off.value == total
            names.append(buf.value.decode())
            assert sd[names[-1]].numel() == n.value
            total += n.value
        assert names == [k for k in fx['keys'] if not k.endswith('num_batches_tracked')]
        assert lib.dsr_gen_state_numel(plan) == total
        assert lib.dsr_gen_workspace_bytes(plan) > 0
    finally:
        lib.dsr_gen_plan_destroy(plan)


def test_generator_unsupported_configurations():
    import ctypes as C
    import torch
    import dsr_b200
    from dsr_b200._lib import lib
    plan = C.c_void_p()
    assert lib.dsr_gen_plan_create(C.byref(plan), 4, 16, 1, 24, 24) == -5      # the reference builds x8 / x16 only
    assert lib.dsr_gen_plan_create(C.byref(plan), 8, 16, 0, 24, 24) == -1
    with pytest.raises(NotImplementedError):
        dsr_b200.Generator(2)
    g = dsr_b200.Generator(8)
    with pytest.raises(RuntimeError):
        g(torch.rand(1, 3, 16, 16))                       # training mode, CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        g.eval()(torch.rand(1, 3, 16, 16))                # CPU tensor: no fallback
    from models.GAN.generator import Generator as DropIn   # drop-in module name (eval_GAN.py:11)
    assert DropIn is dsr_b200.Generator


def test_gan_training_surface_without_gpu():
    """Host logic of the SRGAN training step that needs no GPU: trainer creation, parameter layout == the drop-in
    modules' named_parameters(), same-seed initial state == the oracle's (i.e. the reference's), unsupported
    configurations and CPU tensors raise."""
    import ctypes as C
    import torch
    import dsr_b200
    from dsr_b200 import gan_train as GT
    from dsr_b200._lib import lib
    from oracle import gan_oracle as GO, gan_train_oracle as O
    h = C.c_void_p()
    assert lib.dsr_gant_create(C.byref(h), 9, 24, 24, 8, 16, 1) == -1          # batch <= 8
    assert lib.dsr_gant_create(C.byref(h), 8, 24, 24, 4, 16, 1) == -1          # factor 8 / 16 only
    assert lib.dsr_gant_create(C.byref(h), 8, 40, 40, 8, 16, 1) == -5          # VGG transform only enlarges (patch <= 256)
    assert lib.dsr_gant_create(C.byref(h), 8, 24, 24, 8, 16, 1) == 0
    name = C.create_string_buffer(160)
    off, n = C.c_longlong(), C.c_longlong()
    torch.manual_seed(4)
    G = dsr_b200.Generator(8)
    D = GT.Discriminator((192, 192))
    for net, mod in ((0, G), (1, D)):
        named = list(mod.named_parameters())
        assert lib.dsr_gant_num_params(h, net) == len(named)
        total = 0
        for i, (k, p) in enumerate(named):
            assert lib.dsr_gant_param_info(h, net, i, name, 160, C.byref(off), C.byref(n)) == 0
            assert (name.value.decode(), off.value, n.value) == (k, total, p.numel())
            total += p.numel()
        assert lib.dsr_gant_param_numel(h, net) == total
        bufs = [(k, b) for k, b in mod.named_buffers() if not k.endswith('num_batches_tracked')]
        assert lib.dsr_gant_num_buffers(h, net) == len(bufs)
        for i, (k, b) in enumerate(bufs):
            assert lib.dsr_gant_buffer_info(h, net, i, name, 160, C.byref(off), C.byref(n)) == 0
            assert (name.value.decode(), n.value) == (k, b.numel())
    assert lib.dsr_gant_param_numel(h, 2) == 20024384                            # VGG19 features[:36]
    assert lib.dsr_gant_workspace_bytes(h) > 0
    assert lib.dsr_gant_g_forward(h, None, None, None, None, 1, None) == -1     # unbound / null arguments
    assert lib.dsr_gant_vgg_real(h, None, None) == -1 and lib.dsr_gant_vgg_loss(h, None, None, None, 0, None, None) == -1
    assert lib.dsr_gant_dense_grad_overwrite(None, 1) == -1 and lib.dsr_gant_dense_grad_overwrite(h, 1) == 0
    assert lib.dsr_gant_dense_grad_overwrite(h, 0) == 0
    lib.dsr_gant_destroy(h)
    # image writer path: argument checks happen before any CUDA call
    assert lib.dsr_image_to_u8_hwc(None, 1, 3, 8, 8, 0, None, None) == -1
    with pytest.raises(RuntimeError):
        dsr_b200.to_uint8_hwc(torch.rand(3, 8, 8))                               # CPU tensor: no fallback
    # same seed -> the reference's initial state (incl. the running statistics fc_input_shape leaves behind)
    torch.manual_seed(4)
    sdG = GO.init_state_dict(8)
    sdD = O.init_discriminator((192, 192))
    for k, v in G.state_dict().items():
        assert torch.equal(v, sdG[k]), k
    for k, v in D.state_dict().items():
        assert torch.equal(v, sdD[k]), k
    with pytest.raises(RuntimeError):
        D(torch.rand(2, 3, 192, 192))
    with pytest.raises(RuntimeError):
        GT.Vgg19Loss(pretrained=False)(torch.rand(1, 3, 64, 64), torch.rand(1, 3, 64, 64))
    from models.GAN.discriminator import Discriminator as DropD
    from utils.GAN import PerceptualLoss, get_loss_D                          # noqa: F401
    assert DropD is GT.Discriminator
