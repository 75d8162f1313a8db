"""The image writer path of eval_GAN.py:50-53 / utils/common.py: the drop-in `utils.common` on the CPU, the device-side
uint8 conversion bit for bit against numpy, and the reference's OWN `GAN_ISR_Batch_eval` (unmodified, baseline/_ref)
over the drop-in modules against the mirror `dsr_b200.GAN_ISR_Batch_eval` (PNG files byte-identical)."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
PKG = os.path.join(ROOT, 'deep-super-resolution_b200')


def _load_common():
    spec = importlib.util.spec_from_file_location('dsr_dropin_common', os.path.join(PKG, 'utils', 'common.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_dropin_common_matches_the_reference_functions(tmp_path):
    C = _load_common()
    # the names the reference scripts pick up through `from utils.common import *` (eval_GAN.py uses np without importing it)
    for name in ('np', 'torch', 'os', 'Image', 'datetime', 're', 'OrderedDict', 'save_model', 'save_image', 'save_log',
                 'load_model', 'pil_to_np', 'np_to_pil', 'np_to_torch', 'torch_to_np', 'lpips'):
        assert hasattr(C, name), name
    rng = np.random.default_rng(0)
    img = rng.random((3, 9, 7), dtype=np.float32) * 1.4 - 0.2              # outside [0, 1] on both sides: np_to_pil clips
    pil = C.np_to_pil(img)
    want = np.clip(img * 255, 0, 255).astype(np.uint8).transpose(1, 2, 0)  # utils/common.py:81-86
    assert np.array_equal(np.array(pil), want)
    back = C.pil_to_np(pil)
    assert back.shape == (3, 9, 7) and back.dtype == np.float32 and np.array_equal(back, want.transpose(2, 0, 1) / np.float32(255.))
    t = C.np_to_torch(img)
    assert t.shape == (1, 3, 9, 7) and np.array_equal(C.torch_to_np(t), img)
    C.save_image(want, 'a', str(tmp_path))
    from PIL import Image
    assert np.array_equal(np.array(Image.open(tmp_path / 'images' / 'a.png')), want)
    # save_model / load_model incl. the DataParallel 'module.' prefix (utils/common.py:46-60)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    C.save_model(net, 'm', str(tmp_path))
    net2 = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    C.load_model(net2, str(tmp_path / 'm.pth'))
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    torch.save({'module.' + k: v for k, v in net.state_dict().items()}, tmp_path / 'dp.pth')
    net3 = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    C.load_model(net3, str(tmp_path / 'dp.pth'))
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net3.state_dict().values()))
    C.save_log(str(tmp_path), avg_psnr=1.5, n=2)
    logs = [f for f in os.listdir(tmp_path) if f.endswith('_log.txt')]
    assert len(logs) == 1 and open(tmp_path / logs[0]).read() == 'avg_psnr: 1.5\nn: 2\n'


@pytest.mark.skipif(not os.path.isdir(REF), reason='baseline/_ref not staged (no /root/reference at build time)')
def test_dropin_common_against_the_reference_module():
    spec = importlib.util.spec_from_file_location('ref_common', os.path.join(REF, 'utils', 'common.py'))
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    C = _load_common()
    rng = np.random.default_rng(1)
    img = rng.random((3, 12, 10), dtype=np.float32) * 1.6 - 0.3
    assert np.array_equal(np.array(C.np_to_pil(img)), np.array(R.np_to_pil(img)))
    assert np.array_equal(C.pil_to_np(C.np_to_pil(img)), R.pil_to_np(R.np_to_pil(img)))
    gray = rng.random((1, 5, 6), dtype=np.float32)
    assert np.array_equal(np.array(C.np_to_pil(gray)), np.array(R.np_to_pil(gray)))
    assert np.array_equal(C.pil_to_np(C.np_to_pil(gray)), R.pil_to_np(R.np_to_pil(gray)))


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(1, 3, 64, 48), (2, 3, 37, 21), (1, 1, 16, 20), (3, 64, 64)])
def test_uint8_conversion_is_bit_exact_against_numpy(shape):
    sys.path[:0] = [p for p in (PKG, ROOT) if p not in sys.path]
    import dsr_b200
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(shape, generator=g) * 3.0 - 1.0          # tanh-like negatives and values above 1: the cast wraps
    flat = x.view(-1)
    flat[:6] = torch.tensor([-0.5, 1.0, 254.9 / 255, -1.0, float('nan'), 2.0])
    xn = x.numpy()
    with np.errstate(invalid='ignore'):
        if x.dim() == 3:
            want = (xn.transpose(1, 2, 0) * 255).astype(np.uint8)                                 # eval_GAN.py:52
            want_clip = np.clip(xn * 255, 0, 255).astype(np.uint8).transpose(1, 2, 0)           # utils/common.py:81
        else:
            want = (xn.transpose(0, 2, 3, 1) * 255).astype(np.uint8)
            want_clip = np.clip(xn * 255, 0, 255).astype(np.uint8).transpose(0, 2, 3, 1)
    got = dsr_b200.to_uint8_hwc(x.cuda()).cpu().numpy()
    got_clip = dsr_b200.to_uint8_hwc(x.cuda(), clip=True).cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want)
    assert np.array_equal(got_clip, want_clip)
    with pytest.raises(RuntimeError):
        dsr_b200.to_uint8_hwc(x)                            # no CPU fallback


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'run_reference_gan_eval.py'), *args],
                         capture_output=True, text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.isdir(REF), reason='baseline/_ref not staged (no /root/reference at build time)')
def test_reference_eval_loop_over_its_own_modules_on_cpu():
    r = _run('--impl', 'reference', '--device', 'cpu', '--images', '1', '--lr-size', '8')
    assert r['modules'].startswith('baseline/_ref/') and r['utils_common'].startswith('baseline/_ref/')
    assert len(r['png_sha256']) == 1 and set(r['metrics']) == {'avg_psnr', 'avg_ssim', 'avg_lpips'}


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(REF), reason='baseline/_ref not staged (no /root/reference at build time)')
def test_reference_eval_loop_over_the_drop_in_modules_and_the_mirror():
    r = _run('--impl', 'ours', '--device', 'cuda', '--images', '2', '--lr-size', '20')
    assert r['modules'].startswith('deep-super-resolution_b200/') and r['utils_common'].startswith('deep-super-resolution_b200/')
    assert r['png_sha256'] == r['mirror_png_sha256']        # device-side conversion == numpy's, PNG files byte-identical
    assert r['mirror_metrics']['avg_psnr'] == pytest.approx(r['metrics']['avg_psnr'], abs=1e-5)
    assert r['mirror_metrics']['avg_ssim'] == pytest.approx(r['metrics']['avg_ssim'], abs=1e-6)
    assert r['mirror_metrics']['avg_lpips'] is None
