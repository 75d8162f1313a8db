"""The drop-in claim, executed: the reference's OWN per-image driver (DIP.DIP_ISR, DIP.py:22-123, unmodified, staged in
the git-ignored baseline/_ref by __graft_entry__.build()) runs over this repository's modules when
deep-super-resolution_b200/ is first on sys.path -- BASELINE configs[0] (256x256, 100 iterations)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')


def run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'run_reference_dip.py'), *args],
                         capture_output=True, text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.isdir(REF), reason='baseline/_ref not staged (no /root/reference at build time)')
def test_reference_driver_over_its_own_modules_on_cpu():
    r = run('--impl', 'reference', '--device', 'cpu', '--size', '64', '--iters', '4', '--log-freq', '2')
    assert r['modules'].startswith('baseline/_ref/') and r['resolved_shape'] == [1, 3, 64, 64] and len(r['psnrs']) == 2


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(REF), reason='baseline/_ref not staged (no /root/reference at build time)')
def test_reference_driver_over_the_drop_in_modules():
    r = run('--impl', 'ours', '--device', 'cuda', '--size', '256', '--iters', '100', '--log-freq', '25')
    print(f"reference DIP.DIP_ISR over dsr_b200 at 256^2: {r['it_per_s']:.1f} it/s incl. set-up; PSNR {r['psnrs']}")
    assert r['modules'].startswith('deep-super-resolution_b200/')
    assert r['resolved_shape'] == [1, 3, 256, 256] and len(r['psnrs']) == 4 and len(r['ssims']) == 4
    assert r['psnrs'][-1] > r['psnrs'][0] + 0.5 and r['ssims'][-1] > r['ssims'][0]      # the fit improves (logged at it 0 .. 75)
    assert r['final_psnr'] > r['psnrs'][0]
