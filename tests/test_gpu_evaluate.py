"""The reference's OWN ``evaluate.py`` running unchanged on top of the drop-in modules (``robustcap_b200/dropin``): a synthetic
AIST++-style ``test.pt`` (schema: ``preprocess.py:50, 289``) + random-init ``best_weights.pt`` + ``not_aligned.txt`` are written
into a scratch asset tree, ``evaluate.evaluate_aist_ours`` (``evaluate.py:20-117``) is executed, and the ``result.pt`` it saves is
compared with the CPU oracle run frame by frame on the same rows.  ``evaluate.py`` itself comes from the reference checkout
(``/root/reference`` in the build container; on the GPU box from the archive ``__graft_entry__.build()`` leaves in the git-ignored
``baseline/_ref/``), never from this repository.  Also runs the batched equivalent (``pipeline.load_aist_style`` ->
``Net.forward_offline`` -> ``metrics.cal_mpjpe``) and checks it against the same result.
"""
import os
import shutil
import subprocess
import sys
import tarfile

import numpy as np
import pytest
import torch

from robustcap_b200 import synthetic
from test_oracle_golden import pose_angle

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCHIVE = os.path.join(REPO, 'baseline', '_ref', 'reference_src.tar.gz')


def reference_dir(tmp_path):
    if os.path.isdir('/root/reference'):
        return '/root/reference'
    if os.path.exists(ARCHIVE):
        with tarfile.open(ARCHIVE) as tf:
            tf.extractall(tmp_path)
        return str(tmp_path / 'reference')
    return None


def make_dataset(root, n_seq=1, T=20, seed=5):
    """Synthetic AIST++-style test.pt: per sequence axis-angle pose / tran, 9 cameras (K, world->camera T), MediaPipe key points in
    normalised image coordinates with confidence, world-frame IMU orientations / accelerations."""
    g = torch.Generator().manual_seed(seed)
    d = {k: [] for k in ('name', 'pose', 'tran', 'joint2d_mp', 'cam_K', 'cam_T', 'imu_ori', 'imu_acc')}
    for s in range(n_seq):
        inp = synthetic.make_inputs(1, T, seed=seed + s, conf='mixed')
        d['name'].append('gBR_sBM_cAll_d04_mBR0_ch%02d' % s)
        d['pose'].append(0.3 * torch.randn(T, 72, generator=g))
        d['tran'].append(torch.randn(T, 3, generator=g) * 0.1)
        K = torch.tensor([[1500., 0, 960], [0, 1500., 540], [0, 0, 1]]).expand(9, 3, 3).clone()
        Tcw = torch.eye(4).expand(9, 4, 4).clone()
        Tcw[:, :3, :3] = synthetic._random_rotations(9, g)
        Tcw[:, :3, 3] = torch.tensor([0., 0., 4.]) + 0.2 * torch.randn(9, 3, generator=g)
        kp = torch.rand(9, T, 33, 3, generator=g)
        kp[..., 2] = (0.6 + 0.4 * torch.rand(9, T, 1, generator=g)).expand(9, T, 33)
        d['joint2d_mp'].append(kp)
        d['cam_K'].append(K)
        d['cam_T'].append(Tcw)
        d['imu_ori'].append(inp['oric'][0])
        d['imu_acc'].append(inp['accc'][0])
    os.makedirs(os.path.join(root, 'data', 'dataset_work', 'AIST'), exist_ok=True)
    torch.save(d, os.path.join(root, 'data', 'dataset_work', 'AIST', 'test.pt'))
    with open(os.path.join(root, 'data', 'dataset_work', 'AIST', 'not_aligned.txt'), 'w') as f:
        f.write(d['name'][0].replace('cAll', 'c03') + '\n')
    os.makedirs(os.path.join(root, 'data', 'weights', 'sig_mp'), exist_ok=True)
    torch.save(synthetic.make_state_dict(0, 'contact'), os.path.join(root, 'data', 'weights', 'sig_mp', 'best_weights.pt'))
    return d


@pytest.mark.parametrize('run_smplify', [False, True])
def test_reference_evaluate_runs_on_dropin(tmp_path, run_smplify):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    ref = reference_dir(tmp_path)
    if ref is None:
        pytest.skip('no reference checkout / archive (run __graft_entry__.build() in the build container)')
    from robustcap_b200 import _lib
    _lib.build()
    root = str(tmp_path / 'assets')
    synthetic.write_assets(root, 0)
    d = make_dataset(root)
    code = r'''
import sys, warnings
warnings.filterwarnings('ignore')
sys.path[:0] = [%r, %r, %r]
import evaluate
assert evaluate.__file__.startswith(%r), evaluate.__file__
evaluate.evaluate_aist_ours(run_smplify=%r)
print('evaluate done')
''' % (os.path.join(REPO, 'robustcap_b200', 'dropin'), REPO, ref, ref, run_smplify)
    out = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and 'evaluate done' in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
    assert 'mpjpe, pve, pmpjpe:' in out.stdout and 'absolute root position error:' in out.stdout
    pose_p, pose_t, tran_p, tran_t = torch.load(os.path.join(root, 'data', 'dataset_work', 'AIST', 'result.pt'))
    assert len(pose_p) == 9 and pose_p[0].shape == (20, 24, 3, 3) and tran_p[0].shape == (20, 3)
    assert all(torch.isfinite(p).all() for p in pose_p)
    if run_smplify:
        return
    # ---- the same rows through the CPU oracle (per-frame loop, evaluate.py:68-85) and through the batched drop-in flow ----
    from oracle.kinematics import BodyOracle
    from oracle.fusion import FusionOracle
    import robustcap_b200 as rb
    from robustcap_b200 import pipeline, metrics
    assets = synthetic.write_assets(root, 0)
    sd = synthetic.make_state_dict(0, 'contact')
    orc = FusionOracle(sd, BodyOracle(assets['smpl_file']))
    body = rb.ParametricModel(assets['smpl_file'])
    batch, valid = pipeline.load_aist_style(d, not_aligned={d['name'][0].replace('cAll', 'c03')})
    assert valid == [0, 1, 3, 4, 5, 6, 7, 8]
    net = rb.Net(body)
    net.load_state_dict(sd)
    bp, bt = net.forward_offline(batch.j2dc, batch.accc, batch.oric, first_tran=batch.first_tran, lengths=batch.lengths, gravity=batch.gravity)
    for cam in (0, 4, 8):
        Tcw = d['cam_T'][0][cam]
        Kinv = d['cam_K'][0][cam].inverse()
        oric = Tcw[:3, :3].matmul(d['imu_ori'][0])
        accc = Tcw.matmul(torch.cat((d['imu_acc'][0], torch.zeros(20, 6, 1)), dim=-1).unsqueeze(-1)).squeeze(-1)[..., :3]
        j2 = d['joint2d_mp'][0][cam].clone()
        px = torch.stack((j2[..., 0] * 1920, j2[..., 1] * 1080, torch.ones(20, 33)), dim=-1)
        j2dc = Kinv.matmul(px.unsqueeze(-1)).squeeze(-1)
        j2dc[..., 2] = j2[..., 2]
        first_tran = Tcw.matmul(torch.cat((d['tran'][0], torch.ones(20, 1)), dim=-1).unsqueeze(-1)).squeeze(-1)[0, :3]
        op, ot = orc.run(j2dc, accc, oric, first_tran=first_tran, gravity=Tcw[:3, :3].mm(torch.tensor([0, -1, 0.]).view(3, 1)).view(3))
        assert pose_angle(pose_p[cam], op).max().item() < 1e-4 and (tran_p[cam] - ot).abs().max().item() < 1e-3, cam
        assert pose_angle(bp[cam].cpu(), op).max().item() < 1e-4 and (bt[cam].cpu() - ot).abs().max().item() < 1e-3, cam
    err = metrics.cal_mpjpe(body, torch.from_numpy(np.load(assets['j_regressor'])).float(), bp[0], pose_t[0].cuda(), cal_pampjpe=True)
    assert err.shape == (3,) and torch.isfinite(err).all()
