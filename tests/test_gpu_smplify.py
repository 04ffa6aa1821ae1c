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


@pytest.mark.parametrize('name,max_iter', [('it5', 5), ('it20', 20)])
def test_native_optimizer_within_reference_spread(env, golden_dir, assets, name, max_iter):
    """The device-resident L-BFGS (rc_smplify_run: torch/optim/lbfgs.py's step + strong Wolfe, one thread block per sequence) against
    the reference's result, in rad / m — held to the reference's OWN spread: tests/golden/smplify_spread_*.npz holds the result of the
    identical reference call under 1 / 2 / 8 CPU threads and with the start point perturbed by one float32 rounding (2e-7 relative).
    Measured there: the reference moves by 3.6e-3 (rotation-matrix entries) / 2.5 mm at max_iter = 5 and by 0.12 / 13 mm at
    max_iter = 20 under that noise (the line search takes discrete decisions on float32 sums of ~1e5 terms), so no implementation can
    sit closer than that to "the" reference result.  Also asserted: the first closure value reproduces the reference's loss to 1e-5,
    the objective decreases, and the native optimiser ends at least as low as torch.optim.LBFGS driving the same closure."""
    rb, smplify, body = env
    from oracle.kinematics import BodyOracle
    from oracle.smplify import SmplifyOracle
    from test_oracle_smplify import objective
    g = load(golden_dir, 'smplify_%s.npz' % name)
    sp = load(golden_dir, 'smplify_spread_%s.npz' % name)
    T = g['pose_in'].shape[0]
    ref_p, ref_t = g['pose_out'], g['tran_out']
    variants = [(sp['pose_out_%s' % k], sp['tran_out_%s' % k]) for k in ('t1', 't2', 't8', 'eps0', 'eps1', 'eps2')]
    spread_rad = max(pose_angle(p, ref_p).max().item() for p, _ in variants)
    spread_m = max((t - ref_t).abs().max().item() for _, t in variants)
    res = {}
    for opt in ('native', 'torch'):
        smplify.TemporalSMPLify.optimizer = opt
        try:
            pose, tran, upd = smplify.smplify_runner(g['pose_in'], g['tran_in'], g['j2d_pix'].clone(), g['imu_ori'], batch_size=T, lr=1e-3,
                                                     use_lbfgs=True, opt_steps=1, cam_k=g['cam_k'], loss_threshold=1e12, max_iter=max_iter)
        finally:
            smplify.TemporalSMPLify.optimizer = 'native'
        res[opt] = (pose.reshape(T, 24, 3, 3), tran)
    checker = SmplifyOracle(BodyOracle(assets['smpl_file']), gmm_file(assets), g['cam_k'], g['imu_ori'], step_size=1e-3, max_iter=max_iter)
    f0 = objective(checker, g, g['pose_in'], g['tran_in'])
    f_ref = objective(checker, g, ref_p, ref_t)
    f = {k: objective(checker, g, v[0], v[1]) for k, v in res.items()}
    d_rad = pose_angle(res['native'][0], ref_p).max().item()
    d_m = (res['native'][1] - ref_t).abs().max().item()
    print('%s: native vs reference result %.3e rad %.3e m | reference spread (threads, 1-ulp start noise) %.3e rad %.3e m | objective start %.1f, '
          'reference %.1f, native %.1f, torch.optim.LBFGS on the native closure %.1f' % (name, d_rad, d_m, spread_rad, spread_m, f0, f_ref, f['native'], f['torch']))
    assert f['native'] < f0 and f['native'] <= 1.01 * f_ref
    assert d_rad <= 1.5 * spread_rad + 1e-4 and d_m <= 1.5 * spread_m + 1e-3


def test_batched_sequences(env, golden_dir):
    """smplify_runner_batch: S sequences in one launch (one thread block each) == the same sequences one at a time, bit for bit."""
    rb, smplify, body = env
    g = load(golden_dir, 'smplify_it5.npz')
    T = g['pose_in'].shape[0]
    S = 5
    gen = torch.Generator().manual_seed(3)
    pose = g['pose_in'].unsqueeze(0).repeat(S, 1, 1, 1, 1)
    tran = g['tran_in'].unsqueeze(0).repeat(S, 1, 1) + 0.01 * torch.randn(S, 1, 3, generator=gen)
    kp = g['j2d_pix'].unsqueeze(0).repeat(S, 1, 1, 1)
    kp[..., :2] += torch.randn(S, T, 33, 2, generator=gen)
    ori = g['imu_ori'].unsqueeze(0).repeat(S, 1, 1, 1, 1)
    bp, bt, stats = smplify.smplify_runner_batch(pose, tran, kp, ori, g['cam_k'], lr=1e-3, max_iter=5)
    assert bp.shape == (S, T, 24, 3, 3) and bt.shape == (S, T, 3) and stats.shape == (S, 4)
    print('batched SMPLify stats (first loss, final loss, evaluations, iterations):', stats.cpu().tolist())
    assert (stats[:, 1] < stats[:, 0]).all() and (stats[:, 2] <= 7).all()            # max_eval = 6; torch's line search may run one evaluation past it
    for s_ in (0, S - 1):
        p1, t1, _ = smplify.smplify_runner(pose[s_], tran[s_], kp[s_].clone(), ori[s_], batch_size=T, lr=1e-3, cam_k=g['cam_k'],
                                           loss_threshold=1e12, max_iter=5)
        assert torch.equal(p1, bp[s_].cpu()) and torch.equal(t1, bt[s_].cpu())


def test_shaped_body(env, golden_dir):
    """``smplify_runner(..., shape=betas)`` (run.py:21, temporal_smplify.py:84-86, 158-159): the closure skins the SHAPED rest pose.
    The native constants of ``ParametricModel.shaped(betas)`` must give the same 33 points as the full shaped mesh path (golden-pinned
    in test_smpl_forward_kinematics), and the runner must reduce its objective on the shaped body."""
    rb, smplify, body = env
    from robustcap_b200.net import sync_mp3d
    g = load(golden_dir, 'smplify_it5.npz')
    k = load(golden_dir, 'kinematics.npz')
    shape = k['shape'][0]
    T = g['pose_in'].shape[0]
    sb = body.shaped(shape)
    _, kp = sb.keypoints33(g['pose_in'].cuda(), g['tran_in'].cuda())
    _, gj, gv = body.forward_kinematics(g['pose_in'].cuda(), shape=shape.cuda().expand(T, 10), tran=g['tran_in'].cuda(), calc_mesh=True)
    for t in (0, T - 1):
        assert (kp[t] - sync_mp3d(gv[t], gj[t])).abs().max().item() < 5e-6
    _, kp_mean = body.keypoints33(g['pose_in'].cuda(), g['tran_in'].cuda())
    assert (kp - kp_mean).abs().max().item() > 1e-3                        # the shape really moves the points
    sm = smplify.TemporalSMPLify(cam_k=g['cam_k'], imu_ori=g['imu_ori'], step_size=1e-3, batch_size=T, max_iter=5, shape=shape)
    pose, tran, rl = sm(g['pose_in'], g['tran_in'], g['j2d_pix'].clone())
    st = sm.last_stats.cpu()
    assert st[0, 1] < st[0, 0] and torch.isfinite(pose).all()
    pose2, tran2, upd = smplify.smplify_runner(g['pose_in'], g['tran_in'], g['j2d_pix'].clone(), g['imu_ori'], batch_size=T, lr=1e-3,
                                               cam_k=g['cam_k'], loss_threshold=1e12, shape=shape, max_iter=5)
    assert torch.equal(pose2.reshape(-1), pose.reshape(-1).cpu())
