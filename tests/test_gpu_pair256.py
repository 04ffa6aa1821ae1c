"""Opt-in 256 x 256 CTA-pair variant of the grouped kernel (RC_PH_PAIR=2, csrc/phase_tc.cu) against the default 128 x 128 kernel on the
same batch.  The variant is selected per process, so both run in child processes (tests/gpu_pair_ab.py --child).  The variant is NOT
parity-validated at the 1e-4 rad bar (profiles/r03_pair256.md): this test keeps it running and bounds its distance from the default."""
import os
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B', [300])
def test_pair256_variant_close_to_default(B, tmp_path):
    from gpu_pair_ab import angle
    T = 8
    res = {}
    for v in ('0', '2'):
        out = str(tmp_path / ('pair_%s.pt' % v))
        env = dict(os.environ, RC_PH_PAIR=v)
        r = subprocess.run([sys.executable, os.path.join(HERE, 'gpu_pair_ab.py'), '--child', str(B), str(T), 'mixed', out],
                           env=env, timeout=600, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        res[v] = torch.load(out)
    an = angle(res['2']['pose'], res['0']['pose'])
    q = an.flatten().kthvalue(int(an.numel() * 0.999)).values.item()
    dt = (res['2']['tran'] - res['0']['tran']).abs().max().item()
    print('pair256 vs default (%d x %d): pose max %.2e rad, 99.9 %% %.2e rad, tran max %.2e m' % (B, T, an.max().item(), q, dt))
    assert bool(torch.isfinite(res['2']['pose']).all())
    assert q < 5e-5 and dt < 1e-4            # two fp32-accurate evaluations: reduction-order noise (same bound as the SIMT comparison)
