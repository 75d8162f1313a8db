"""The generator oracle (oracle/gan_oracle.py) against the fixtures recorded from the unmodified reference
(oracle/make_golden_gan.py): same-seed initialisation checksums and eval-mode outputs."""
import pytest
import torch

from oracle import gan_oracle as go

CASES = ['gan_f8_2x20x24.pt', 'gan_f8_1x17x23.pt', 'gan_f16_1x16x16.pt']


def checksum(t):
    t = t.detach().double().flatten()
    return (float(t.sum()), float(t.abs().sum()))


def state_for(fx):
    torch.manual_seed(fx['seed'])
    sd = go.init_state_dict(fx['factor'])
    go.perturb_trained_state(sd, fx['perturb_seed'])
    return sd


@pytest.mark.parametrize('name', CASES)
def test_generator_init_matches_reference(golden, name):
    fx = golden(name)
    sd = state_for(fx)
    assert list(sd.keys()) == fx['keys']                      # state_dict order of the reference module
    for k, v in sd.items():
        assert checksum(v) == pytest.approx(fx['checksums'][k], rel=1e-12, abs=1e-12), k


@pytest.mark.parametrize('name', CASES)
def test_generator_forward_matches_reference(golden, name):
    fx = golden(name)
    torch.set_num_threads(1)
    y = go.generator_forward(state_for(fx), fx['x'], fx['factor'])
    assert y.shape == fx['y'].shape
    assert float((y - fx['y']).abs().max()) <= 1e-6


def test_generator_flops_kat():
    # 98.13 GFLOP per 96 x 96 image at factor 8
    assert go.flops_per_image(96, 96, 8) == pytest.approx(98.13e9, rel=1e-3)
