import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_oracle_golden import load, pose_angle
from test_oracle_smplify import objective, gmm_file
import robustcap_b200 as rb
from robustcap_b200 import _lib, smplify, synthetic
from oracle.kinematics import BodyOracle
from oracle.smplify import SmplifyOracle
_lib.build()
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
body = rb.ParametricModel(assets['smpl_file'])
smplify.TemporalSMPLify.body_model = body
os.chdir(os.path.dirname(os.path.dirname(assets['gmm_dir'])))
for name, mi in (('it5', 5), ('it20', 20)):
    g = load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'), 'smplify_%s.npz' % name)
    T = g['pose_in'].shape[0]
    sm = smplify.TemporalSMPLify(cam_k=g['cam_k'], imu_ori=g['imu_ori'], step_size=1e-3, batch_size=T, max_iter=mi)
    losses = []
    orig = sm._native
    def spy(*a, **k):
        r = orig(*a, **k)
        if r[0] is not None: losses.append(r[0].item())
        return r
    sm._native = spy
    pose, tran, rl = sm(g['pose_in'].reshape(T, -1), g['tran_in'], g['j2d_pix'].clone())
    chk = SmplifyOracle(BodyOracle(assets['smpl_file']), gmm_file(assets), g['cam_k'], g['imu_ori'], step_size=1e-3, max_iter=mi)
    f0 = objective(chk, g, g['pose_in'], g['tran_in']); fr = objective(chk, g, g['pose_out'], g['tran_out'])
    fn = objective(chk, g, pose.cpu().reshape(T, 24, 3, 3), tran.cpu())
    print(name, 'evals', sm.evals, 'losses', ['%.1f' % l for l in losses])
    print('   f0 %.1f f_ref %.1f f_new %.1f  pose-vs-ref %.2e moved %.2e tran diff %.2e' % (f0, fr, fn,
          pose_angle(pose.cpu().reshape(-1, 3, 3), g['pose_out'].reshape(-1, 3, 3)).max().item(),
          pose_angle(g['pose_out'].reshape(-1, 3, 3), g['pose_in'].reshape(-1, 3, 3)).max().item(), (tran.cpu() - g['tran_out']).abs().max().item()))
