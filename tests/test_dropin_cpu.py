"""The reference's own evaluate.py must import unchanged against the drop-in modules (robustcap_b200/dropin first on sys.path).
Needs the reference checkout, i.e. runs in the build container only (skipped on the GPU box)."""
import os
import subprocess
import sys

import pytest

REF = '/root/reference'
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')
def test_evaluate_imports_against_dropin(assets):
    root = os.path.dirname(os.path.dirname(assets['gmm_dir']))
    code = r'''
import sys, warnings
warnings.filterwarnings('ignore')
sys.path[:0] = [%r, %r, %r]
import evaluate
assert evaluate.__file__.startswith(%r)
assert evaluate.art.ParametricModel.__module__ == 'robustcap_b200.model'
assert evaluate.art.math.r6d_to_rotation_matrix.__module__ == 'robustcap_b200.math'
assert evaluate.smplify_runner.__module__ == 'robustcap_b200.smplify'
assert evaluate.RNN.__module__ == 'robustcap_b200.rnn'
from net.sig_mp import Net
assert Net.__module__ == 'robustcap_b200.net'
import utils
assert utils.reconstruction_error.__module__ == 'robustcap_b200.metrics'
sd = Net(evaluate.body_model).state_dict()
assert len(sd) == 6 * 12 + 6 and sum(v.numel() for v in sd.values()) == 63424546      # key set / size of SURVEY.md section 5
print('ok')
''' % (os.path.join(REPO, 'robustcap_b200', 'dropin'), REPO, REF, REF)
    out = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr[-2000:]
