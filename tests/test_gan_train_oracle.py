"""The SRGAN training oracle (oracle/gan_train_oracle.py) against outputs of the UNMODIFIED reference's do_epoch
(train_GAN.py:38-71), recorded by oracle/make_golden_gan_train.py.  CPU only."""
import pytest
import torch

from oracle import gan_oracle as GO
from oracle import gan_train_oracle as O

CASES = ['gan_train_2x8x8.pt', 'gan_train_3x16x8.pt']


def _checksum(t):
    t = t.detach().double().flatten()
    return (float(t.sum()), float(t.abs().sum()))


def _sample(t, n=64):
    f = t.detach().flatten()
    step = max(1, f.numel() // n)
    return f[::step][:n]


def build_state(fx):
    """Same-seed initial state of G, D (construction order of make_golden_gan_train.py) and the random VGG."""
    torch.manual_seed(fx['seed'])
    sdG = GO.init_state_dict(fx['factor'])
    h, w = fx['lr_hw']
    sdD = O.init_discriminator((h * fx['factor'], w * fx['factor']))
    torch.manual_seed(fx['vgg_seed'])
    sdV = O.init_vgg()
    return sdG, sdD, sdV


@pytest.mark.parametrize('case', CASES)
def test_do_epoch_reproduces_reference(golden, case):
    fx = golden(case)
    torch.set_num_threads(1)
    sdG, sdD, sdV = build_state(fx)
    for net, sd in (('G', sdG), ('D', sdD)):
        for k, c in fx['init'][net].items():
            assert _checksum(sd[k]) == pytest.approx(c, rel=1e-12, abs=1e-12), (net, k)
    out = O.do_epoch(sdG, sdD, sdV, fx['LR'], fx['HR'], fx['lr'], fx['factor'])
    assert float(out['loss_D']) == pytest.approx(fx['loss_D'], rel=1e-5)
    assert float(out['loss_G']) == pytest.approx(fx['loss_G'], rel=1e-5)
    # generator gradient (what loss_G.backward() left in .grad): every tensor, checksum and strided sample.  The
    # restated transform (two interpolation matrices) and ATen's resize kernel differ in the last bits; 16 random-weight
    # ReLU layers turn that into ~1e-4 relative on the gradient -- hence 3e-3 here while the losses agree to 1e-5
    for k, (c, s) in fx['grad_G'].items():
        g = out['gG'][k]
        scale = max(c[1], 1e-12)
        assert abs(_checksum(g)[1] - c[1]) <= 3e-3 * scale, k
        assert torch.allclose(_sample(g), s, rtol=2e-2, atol=2e-2 * scale / max(g.numel(), 1)), k
    # the fixture run logs metrics in its only epoch (train_GAN.py:104-112): one more train-mode generator pass, under
    # no_grad and with the updated weights, which moves the generator's running statistics a third time
    with torch.no_grad():
        ns = {}
        O.generator_train(sdG, fx['LR'], fx['factor'], 16, ns)
        sdG.update(ns)
    # post-step state: parameters after the Adam steps, running statistics after the 2 (G) / 3 (D) train-mode passes
    for net, sd in (('G', sdG), ('D', sdD)):
        for k, c in fx['post'][net].items():
            if k.endswith('num_batches_tracked'):
                continue
            assert _checksum(sd[k])[1] == pytest.approx(c[1], rel=2e-5, abs=1e-7), (net, k)
            assert torch.allclose(_sample(sd[k]).float(), fx['post_sample'][net][k].float(), rtol=1e-4,
                                  atol=2.5 * fx['lr']), (net, k)


def test_vgg_transform_matches_torchvision():
    """The restated transform against torchvision's own preset object (utils/GAN.py:76-77)."""
    from torchvision.models import VGG19_Weights
    g = torch.Generator().manual_seed(3)
    for shape in ((2, 3, 64, 64), (1, 3, 192, 192), (1, 3, 128, 64), (1, 3, 48, 96)):
        x = torch.rand(shape, generator=g)
        ref = VGG19_Weights.IMAGENET1K_V1.transforms()(x)
        assert torch.allclose(O.vgg_transform(x), ref, atol=1e-6), shape
        assert torch.allclose(O.vgg_transform_explicit(x), ref, atol=3e-6), shape
