"""GPU parity of the native SMPLify objective / gradient / runner against the reference's golden outputs (run with -m gpu)."""
import os

import pytest
import torch

from test_oracle_golden import load, pose_angle
from test_oracle_smplify import check_optimum, gmm_file

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env(assets):
    import robustcap_b200 as rb
    from robustcap_b200 import _lib, smplify
    _lib.build()
    body = rb.ParametricModel(assets['smpl_file'])
    smplify.TemporalSMPLify.body_model = body
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.dirname(assets['gmm_dir'])))      # the reference resolves 'data/dataset_work' relative to cwd
    yield rb, smplify, body
    os.chdir(cwd)


@pytest.mark.parametrize('name,max_iter', [('it5', 5), ('it20', 20)])
def test_objective_and_gradient(env, golden_dir, assets, name, max_iter):
    rb, smplify, body = env
    g = load(golden_dir, 'smplify_%s.npz' % name)
    T = g['pose_in'].shape[0]
    sm = smplify.TemporalSMPLify(cam_k=g['cam_k'], imu_ori=g['imu_ori'], step_size=1e-3, batch_size=T, max_iter=max_iter)
    kp = g['j2d_pix'].clone()
    j2d, conf = sm._prep_keypoints(kp)
    assert (kp[:, sm.ign_mp_joints, 2] == 0).all()                    # zeroed in place like the reference
    aa = rb.math.rotation_matrix_to_axis_angle(g['pose_in'].cuda()).reshape(T, 72).contiguous()
    assert (aa.cpu() - g['aa_init']).abs().max() < 2e-6
    _, ref3d = body.keypoints33(g['pose_in'].cuda(), g['tran_in'].cuda())
    loss, gp, gt, _ = sm._native(g['aa_init'].cuda().contiguous(), g['tran_in'].cuda().contiguous(), j2d, conf, ref3d.contiguous(), 0, True)
    rel = abs(loss.item() - g['loss_init'].item()) / abs(g['loss_init'].item())
    gerr = (gp.cpu() - g['grad_pose']).abs().max().item() / g['grad_pose'].abs().max().item()
    terr = (gt.cpu() - g['grad_tran']).abs().max().item() / g['grad_tran'].abs().max().item()
    print('loss rel err %.2e, grad pose rel err %.2e, grad tran rel err %.2e' % (rel, gerr, terr))
    assert rel < 1e-5 and gerr < 2e-4 and terr < 2e-4
    rl = sm.get_fitting_loss(g['pose_in'], g['tran_in'], g['j2d_pix'].clone())
    assert (rl.cpu() - g['reproj_init']).abs().max() <= 1e-3 * g['reproj_init'].abs().max()


@pytest.mark.parametrize('name,max_iter', [('it5', 5), ('it20', 20)])
def test_runner(env, golden_dir, assets, name, max_iter):
    rb, smplify, body = env
    from oracle.kinematics import BodyOracle
    from oracle.smplify import SmplifyOracle
    g = load(golden_dir, 'smplify_%s.npz' % name)
    T = g['pose_in'].shape[0]
    pose, tran, upd = smplify.smplify_runner(g['pose_in'], g['tran_in'], g['j2d_pix'].clone(), g['imu_ori'], batch_size=T, lr=1e-3,
                                             use_lbfgs=True, opt_steps=1, cam_k=g['cam_k'], loss_threshold=1e12, max_iter=max_iter)
    assert pose.device.type == 'cpu' and pose.shape == (T, 24, 3, 3) and tran.shape == (T, 3) and upd.shape == (T,)
    checker = SmplifyOracle(BodyOracle(assets['smpl_file']), gmm_file(assets), g['cam_k'], g['imu_ori'], step_size=1e-3, max_iter=max_iter)
    # The L-BFGS path is chaotic at float32-noise level (tests/test_oracle_smplify.py; measured on B200: the first three
    # closure evaluations reproduce the reference's losses 637220.0 / 637200.2 / 637065.1 to the digit, the 4th trial step
    # already differs and the native run ends LOWER: 613 746 vs the reference's 626 948 after max_iter=5, 205 953 vs 213 864
    # after max_iter=20).  What is asserted, with the CPU oracle's objective as the judge: the objective is reduced and is
    # not worse than the reference's result by more than 1 %.
    from test_oracle_smplify import objective
    f0 = objective(checker, g, g['pose_in'], g['tran_in'])
    f_ref = objective(checker, g, g['pose_out'], g['tran_out'])
    f_new = objective(checker, g, pose.reshape(T, 24, 3, 3), tran)
    print('objective: start %.1f, reference result %.1f, native result %.1f' % (f0, f_ref, f_new))
    assert f_new < f0 and f_new <= 1.01 * f_ref
    R = pose.reshape(-1, 3, 3)
    assert (R.transpose(1, 2) @ R - torch.eye(3)).abs().max() < 1e-5
    if max_iter == 20:
        p3, t3, u3 = smplify.smplify_runner(g['pose_in'], g['tran_in'], g['j2d_pix'].clone(), g['imu_ori'], batch_size=T, lr=1e-3,
                                            cam_k=g['cam_k'], loss_threshold=1e-3)
        assert u3 is None and torch.equal(p3, g['pose_in'])
