"""Pin the SMPLify oracle (oracle/smplify.py) against the reference outputs in tests/golden/smplify_*.npz. CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import rotations as R
from oracle.kinematics import BodyOracle
from oracle.smplify import SmplifyOracle, smplify_runner
from test_oracle_golden import load, pose_angle


def objective(sm, g, pose_mat, tran):
    T = pose_mat.shape[0]
    kp = g['j2d_pix'].clone()
    conf = kp[:, :, -1]
    conf[:, sm.ignored] = 0
    aa = R.matrix_to_axis_angle_cv2(pose_mat).reshape(T, -1)
    with torch.no_grad():
        _, ref = sm.points(g['pose_in'], g['tran_in'])
        return sm.loss(aa, tran, kp[:, :, :2], conf, ref).item()


def check_optimum(sm, g, pose, tran):
    T = g['pose_in'].shape[0]
    f0 = objective(sm, g, g['pose_in'], g['tran_in'])
    f_ref = objective(sm, g, g['pose_out'], g['tran_out'])
    f_new = objective(sm, g, pose.reshape(T, 24, 3, 3), tran)
    assert f_new < f0 and abs(f_new - f_ref) < 0.01 * f_ref, (f0, f_ref, f_new)
    moved = pose_angle(g['pose_out'].reshape(-1, 3, 3), g['pose_in'].reshape(-1, 3, 3)).max().item()
    ang = pose_angle(pose.reshape(-1, 3, 3), g['pose_out'].reshape(-1, 3, 3)).max().item()
    assert ang < max(3e-3, 0.05 * moved), (ang, moved)
    assert (tran - g['tran_out']).abs().max().item() < 5e-3


def gmm_file(assets):
    return os.path.join(assets['gmm_dir'], 'gmm_08.pkl')


@pytest.mark.parametrize('name,max_iter', [('it5', 5), ('it20', 20)])
def test_smplify_oracle(golden_dir, assets, name, max_iter):
    g = load(golden_dir, 'smplify_%s.npz' % name)
    body = BodyOracle(assets['smpl_file'])
    T = g['pose_in'].shape[0]
    sm = SmplifyOracle(body, gmm_file(assets), g['cam_k'], g['imu_ori'], step_size=1e-3, max_iter=max_iter)
    # loss value and autograd gradient at the starting point
    kp = g['j2d_pix'].clone()
    conf = kp[:, :, -1]
    conf[:, sm.ignored] = 0.
    aa = R.matrix_to_axis_angle_cv2(g['pose_in']).reshape(T, -1)
    assert torch.equal(aa, g['aa_init'])
    bp = aa.clone().requires_grad_(True)
    tr = g['tran_in'].clone().requires_grad_(True)
    with torch.no_grad():
        _, ref = sm.points(g['pose_in'], g['tran_in'])
    loss = sm.loss(bp, tr, kp[:, :, :2], conf, ref)
    loss.backward()
    assert abs(loss.item() - g['loss_init'].item()) <= 1e-5 * abs(g['loss_init'].item())
    assert (bp.grad - g['grad_pose']).abs().max() <= 1e-4 * g['grad_pose'].abs().max()
    assert (tr.grad - g['grad_tran']).abs().max() <= 1e-4 * g['grad_tran'].abs().max()
    assert (sm.fitting_loss(g['pose_in'], g['tran_in'], g['j2d_pix'].clone()) - g['reproj_init']).abs().max() <= 1e-3 * g['reproj_init'].abs().max()
    # The optimisation.  The strong-Wolfe line search compares float32 losses of ~6e5 whose last bits are rounding noise, so
    # its step decisions are chaotic ("results maybe slightly different ... due to the randomness of the optimization",
    # reference README.md:27).  Parity of the optimiser is therefore asserted on what it is for: the objective reached must
    # agree within 1 % and the parameters must land within 5 % of the distance moved (3 mrad floor).
    po, to, rl = sm(g['pose_in'].reshape(T, -1), g['tran_in'], g['j2d_pix'].clone())
    check_optimum(sm, g, po, to)
    if max_iter == 20:
        p2, t2, upd = smplify_runner(body, gmm_file(assets), g['pose_in'], g['tran_in'], g['j2d_pix'].clone(), g['imu_ori'], T,
                                     g['cam_k'], lr=1e-3, loss_threshold=1e12)
        assert pose_angle(p2.view(-1, 3, 3), g['runner_pose'].reshape(-1, 3, 3)).max().item() < 0.05
        assert (upd == g['runner_update']).float().mean() > 0.9
        p3, t3, upd3 = smplify_runner(body, gmm_file(assets), g['pose_in'], g['tran_in'], g['j2d_pix'].clone(), g['imu_ori'], T,
                                      g['cam_k'], lr=1e-3, loss_threshold=1e-3)
        assert upd3 is None and torch.equal(p3.reshape(-1), g['pose_in'].reshape(-1))
