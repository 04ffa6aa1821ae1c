"""Generate the committed golden fixtures by running the UNMODIFIED reference (``/root/reference``) on CPU.

Run in the build container only (the GPU box has no ``/root/reference``):

    python tests/golden/make_golden.py

The reference is imported read-only with the accommodations of SURVEY.md §8(c): stub modules for the
optional viz/training deps, seeded synthetic assets written by ``robustcap_b200.synthetic.write_assets``
and ``chdir`` into that asset tree so the reference's relative paths resolve.  Weights are NOT stored (254 MB):
they are regenerated from ``(seed, variant)`` by ``robustcap_b200.synthetic.make_state_dict`` and loaded into
the reference ``Net`` with ``load_state_dict``; inputs and reference outputs are stored as float32 ``.npz``.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from robustcap_b200 import synthetic  # noqa: E402

REF = '/root/reference'
ASSET_SEED = 0


def import_reference():
    for name in ('trimesh', 'pyrender', 'smplx', 'wandb'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['smplx'].SMPL = object
    thop = types.ModuleType('thop')
    thop.clever_format = lambda *a, **k: ''
    sys.modules['thop'] = thop
    root = synthetic.default_asset_root()
    synthetic.write_assets(root, ASSET_SEED)
    os.chdir(root)
    sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings('ignore')
    import articulate as art
    from net.sig_mp import Net
    import net.sig_mp as sig
    from net.smplify.run import smplify_runner
    from net.smplify import temporal_smplify as ts
    import config as rconfig
    import utils as rutils
    return dict(art=art, Net=Net, sig=sig, smplify_runner=smplify_runner, ts=ts, config=rconfig, utils=rutils)


def gen_math(ref, out):
    art = ref['art']
    M = art.math
    g = torch.Generator().manual_seed(11)
    d = {}
    r6d = torch.randn(64, 6, generator=g)
    r6d[0] = 0  # NaN -> 0 path (angular.py:263)
    r6d[1, 3:] = r6d[1, :3] * 2  # degenerate (parallel) pair
    d['r6d'] = r6d
    d['r6d_to_R'] = M.r6d_to_rotation_matrix(r6d)
    R = synthetic._random_rotations(64, g)
    d['R'] = R
    d['R_to_r6d'] = M.rotation_matrix_to_r6d(R)
    aa = torch.randn(64, 3, generator=g)
    aa[0] = 0
    aa[1] = torch.tensor([0.6 * np.pi, 0.8 * np.pi, 0.0])  # |theta| = pi
    aa[2] *= 1e-5
    d['aa'] = aa
    d['aa_to_R'] = M.axis_angle_to_rotation_matrix(aa)
    d['batch_rodrigues'] = ref['ts'].batch_rodrigues(aa)
    # cv2.Rodrigues semantics incl. non-orthonormal inputs (angular.py:236-246)
    Rn = torch.cat((R, d['aa_to_R'], R + 1e-3 * torch.randn(64, 3, 3, generator=g),
                    torch.eye(3).expand(2, 3, 3)))
    d['R_noisy'] = Rn
    d['R_to_aa'] = M.rotation_matrix_to_axis_angle(Rn)
    q = torch.randn(64, 4, generator=g)
    d['q'] = q
    d['q_to_R'] = M.quaternion_to_rotation_matrix(q)
    d['q_to_aa'] = M.quaternion_to_axis_angle(q)
    d['aa_to_q'] = M.axis_angle_to_quaternion(aa)
    q2 = torch.randn(64, 4, generator=g)
    d['q2'] = q2
    d['q_prod'] = M.quaternion_product(q, q2)
    d['q_inv'] = M.quaternion_inverse(q)
    d['q_mean'] = M.quaternion_mean(q)
    d['angle_between'] = M.angle_between(R[:32], R[32:])
    np.savez_compressed(os.path.join(out, 'math.npz'), **{k: v.numpy() for k, v in d.items()})


def gen_kinematics(ref, out):
    art = ref['art']
    bm = art.ParametricModel('models/SMPL_male.pkl')
    g = torch.Generator().manual_seed(12)
    d = {}
    B = 5
    pose = synthetic._random_rotations(B * 24, g).view(B, 24, 3, 3)
    tran = torch.randn(B, 3, generator=g)
    shape = torch.randn(B, 10, generator=g)
    d['pose'], d['tran'], d['shape'] = pose, tran, shape
    j0, v0 = bm.get_zero_pose_joint_and_vertex()
    d['zero_j'], d['zero_v'] = j0, v0
    js, vs = bm.get_zero_pose_joint_and_vertex(shape[:2])
    d['zero_j_shape'], d['zero_v_shape'] = js, vs
    d['bone'] = bm.joint_position_to_bone_vector(j0.unsqueeze(0))
    d['bone_to_joint'] = bm.bone_vector_to_joint_position(d['bone'])
    d['fk_R'] = bm.forward_kinematics_R(pose)
    d['ik_R'] = bm.inverse_kinematics_R(d['fk_R'])
    T_local = art.math.transformation_matrix(pose, d['bone'].expand(B, 24, 3))
    d['T_local'] = T_local
    d['fk_T'] = bm.forward_kinematics_T(T_local)
    d['ik_T'] = bm.inverse_kinematics_T(d['fk_T'])
    d['inv_T'] = art.math.inverse_transformation_matrix(T_local)
    gr, gj = bm.forward_kinematics(pose, tran=tran)
    d['fk_grot'], d['fk_joint'] = gr, gj
    gr, gj, gv = bm.forward_kinematics(pose, tran=tran, calc_mesh=True)
    d['fk_mesh_joint'], d['fk_mesh_vert'] = gj, gv[:2]
    d['fk_mesh_vert_mp'] = ref['utils'].sync_mp3d_from_smpl(gv, gj)
    gr, gj, gv = bm.forward_kinematics(pose[:2], shape=shape[:2], tran=tran[:2], calc_mesh=True)
    d['fk_shape_joint'], d['fk_shape_vert'] = gj, gv
    kp = torch.randn(7, 33, 3, generator=g)
    d['kp'] = kp
    d['bbox_scale'] = ref['sig'].get_bbox_scale(kp)
    d['mp_mask'] = torch.tensor(ref['config'].mp_mask)
    d['ji_mask'] = torch.tensor(ref['config'].ji_mask)
    d['vi_mask'] = torch.tensor(ref['config'].vi_mask)
    d['parent'] = torch.tensor([-1] + bm.parent[1:])
    np.savez_compressed(os.path.join(out, 'kinematics.npz'), **{k: v.numpy() for k, v in d.items()})


# (name, weight seed, variant, conf mode, input seed, start, T)
ONLINE_CASES = [
    ('mixed_ff', 0, 'default', 'mixed', 1, 'first_frame', 48),
    ('mixed_ft', 0, 'default', 'mixed', 2, 'first_tran', 48),
    ('high_none', 0, 'default', 'high', 3, 'none', 32),
    ('mid_ft', 0, 'default', 'mid', 4, 'first_tran', 32),
    ('low_ff', 0, 'default', 'low', 5, 'first_frame', 32),
    ('occl_ft', 0, 'default', 'occluded', 6, 'first_tran', 48),
    ('contact_mixed_ft', 0, 'contact', 'mixed', 7, 'first_tran', 64),
    ('contact_high_none', 0, 'contact', 'high', 8, 'none', 64),
    ('snap_high_ft', 0, 'snap', 'high', 9, 'first_tran', 32),
    ('contact_low_ff', 0, 'contact', 'low', 10, 'first_frame', 32),
]


def run_online(ref, sd, inp, start, record=None, live=False):
    Net = ref['Net']
    Net.live = live                      # class attribute, set BEFORE construction (sig_mp.py:91-93) and read again on every frame (:229, :264)
    net = Net()
    net.load_state_dict(sd)
    net.eval()
    Net.gravityc = inp['gravity'].clone()
    T = inp['j2dc'].shape[1]
    hooks = []
    if record is not None:
        for k in (2, 3, 4, 6, 7, 8):
            def mk(k):
                def hook(mod, args, output):
                    record.append((k, output.detach().clone().flatten()))
                return hook
            hooks.append(getattr(net, 'rnn%d' % k).linear2.register_forward_hook(mk(k)))
    poses, trans = [], []
    for t in range(T):
        kw = {}
        if t == 0 and start == 'first_frame':
            kw['first_frame'] = True
        if t == 0 and start == 'first_tran':
            kw['first_tran'] = torch.tensor([0.0, 0.0, 4.0])
        p, tr = net.forward_online(inp['j2dc'][0, t], inp['accc'][0, t], inp['oric'][0, t], **kw)
        poses.append(p)
        trans.append(tr)
    for h in hooks:
        h.remove()
    Net.live = False
    floor_n = len(net.floor_y)
    net.reset_states()
    return torch.stack(poses), torch.stack(trans), floor_n


LIVE_CASES = [
    ('live_mixed_ft', 0, 'contact', 'mixed', 31, 'first_tran', 72),
    ('live_low_none', 0, 'default', 'low', 32, 'none', 40),
]


def gen_online(ref, out):
    cache = {}
    for name, wseed, variant, conf, iseed, start, T in ONLINE_CASES + LIVE_CASES:
        key = (wseed, variant)
        if key not in cache:
            cache[key] = synthetic.make_state_dict(wseed, variant)
        inp = synthetic.make_inputs(1, T, seed=iseed, conf=conf)
        rec = []
        pose, tran, floor_n = run_online(ref, cache[key], inp, start, rec, live=name.startswith('live_'))
        d = {'j2dc': inp['j2dc'][0], 'accc': inp['accc'][0], 'oric': inp['oric'][0], 'gravity': inp['gravity'],
             'pose': pose, 'tran': tran, 'floor_n': torch.tensor(floor_n)}
        # sub-net outputs in call order for the first 6 frames (debug aid: which net diverged first)
        ncall = min(len(rec), 6 * 8)
        d['rec_net'] = torch.tensor([k for k, _ in rec[:ncall]])
        width = 144
        d['rec_out'] = torch.stack([torch.nn.functional.pad(o, (0, width - o.numel())) for _, o in rec[:ncall]])
        np.savez_compressed(os.path.join(out, 'online_%s.npz' % name), **{k: v.numpy() for k, v in d.items()})
        print('online', name, 'floor samples', floor_n, 'max |tran|', float(tran.abs().max()))


SMPLIFY_CASES = [('it5', 5, 24), ('it20', 20, 16)]


def gen_smplify(ref, out):
    ts = ref['ts']
    art = ref['art']
    sd = synthetic.make_state_dict(0, 'default')
    for name, max_iter, T in SMPLIFY_CASES:
        inp = synthetic.make_inputs(1, T, seed=21, conf='high')
        pose, tran, _ = run_online(ref, sd, inp, 'first_tran')
        cam_k = torch.tensor([[1000.0, 0, 960], [0, 1000, 540], [0, 0, 1]])
        bm = art.ParametricModel('models/SMPL_male.pkl')
        _, joint, vert = bm.forward_kinematics(pose, tran=tran, calc_mesh=True)
        j = ref['utils'].sync_mp3d_from_smpl(vert, joint)
        g = torch.Generator().manual_seed(22)
        uv = (cam_k @ (j / j[..., 2:]).unsqueeze(-1)).squeeze(-1)[..., :2] + 5 * torch.randn(T, 33, 2, generator=g)
        j2d_pix = torch.cat((uv, 0.5 + 0.5 * torch.rand(T, 33, 1, generator=g)), dim=-1)
        imu_ori = inp['oric'][0]
        d = {'pose_in': pose, 'tran_in': tran, 'j2d_pix': j2d_pix.clone(), 'imu_ori': imu_ori, 'cam_k': cam_k}
        # loss value and gradient at the initial point (analytic-gradient check)
        sm = ts.TemporalSMPLify(cam_k=cam_k, imu_ori=imu_ori, step_size=1e-3, num_iters=1, use_lbfgs=True,
                                batch_size=T, max_iter=max_iter)
        kp = j2d_pix.clone()
        joints_conf = kp[:, :, -1]
        joints_conf[:, sm.ign_mp_joints] = 0.
        body_pose = art.math.rotation_matrix_to_axis_angle(pose).reshape(T, -1).clone().requires_grad_(True)
        gt = tran.clone().requires_grad_(True)
        _, j0, v0 = bm.forward_kinematics(pose, tran=tran, calc_mesh=True)
        ref3d = ref['utils'].sync_mp3d_from_smpl(v0, j0).detach()
        R = ts.batch_rodrigues(body_pose.view(-1, 3)).view(T, -1, 3, 3)
        gp, jj, vv = bm.forward_kinematics(R, tran=gt, calc_mesh=True)
        mj = ref['utils'].sync_mp3d_from_smpl(vv, jj)
        from net.smplify.losses import temporal_body_fitting_loss
        loss = temporal_body_fitting_loss(body_pose, mj, kp[:, :, :2], joints_conf, sm.pose_prior, cam_k, ref3d,
                                          imu_ori, gp[:, [ts.joint_mask]])
        loss.backward()
        d['aa_init'] = body_pose.detach()
        d['loss_init'] = loss.detach()
        d['grad_pose'] = body_pose.grad
        d['grad_tran'] = gt.grad
        d['reproj_init'] = sm.get_fitting_loss(pose, tran, j2d_pix.clone())
        # the optimisation itself, exactly as smplify_runner does but with max_iter exposed
        po, to, rl = sm(pose.reshape(T, -1).detach(), tran.detach(), j2d_pix.clone())
        d['pose_out'], d['tran_out'], d['reproj_out'] = po.reshape(T, 24, 3, 3), to, rl
        if max_iter == 20:
            p2, t2, upd = ref['smplify_runner'](pose, tran, j2d_pix.clone(), imu_ori, batch_size=T, lr=1e-3,
                                                use_lbfgs=True, opt_steps=1, cam_k=cam_k, loss_threshold=1e12)
            d['runner_pose'], d['runner_tran'], d['runner_update'] = p2, t2, upd
            p3, t3, upd3 = ref['smplify_runner'](pose, tran, j2d_pix.clone(), imu_ori, batch_size=T, lr=1e-3,
                                                 use_lbfgs=True, opt_steps=1, cam_k=cam_k, loss_threshold=1e-3)
            assert upd3 is None
        np.savez_compressed(os.path.join(out, 'smplify_%s.npz' % name),
                            **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()})
        print('smplify', name, 'loss_init', float(loss), 'moved', float((po.reshape(T, 24, 3, 3) - pose).abs().max()))


def gen_smplify_spread(ref, out):
    """The reference's own run-to-run spread of the SMPLify result (README.md:27 "slightly different ... due to the randomness of the
    optimization"): the identical call under 1, 2 and 8 CPU threads (different float32 reduction orders inside the ATen ops).  The
    L-BFGS line search makes discrete decisions on float32 sums of ~1e5 terms, so the trajectories separate after a few closure
    evaluations; a GPU implementation cannot be closer to "the" reference result than the reference is to itself."""
    ts = ref['ts']
    for name, max_iter, T in SMPLIFY_CASES:
        g = dict(np.load(os.path.join(out, 'smplify_%s.npz' % name)))
        pose, tran = torch.from_numpy(g['pose_in']), torch.from_numpy(g['tran_in'])
        d = {}
        for nt in (1, 2, 8):
            torch.set_num_threads(nt)
            sm = ts.TemporalSMPLify(cam_k=torch.from_numpy(g['cam_k']), imu_ori=torch.from_numpy(g['imu_ori']), step_size=1e-3, num_iters=1,
                                    use_lbfgs=True, batch_size=T, max_iter=max_iter)
            po, to, rl = sm(pose.reshape(T, -1).detach(), tran.detach(), torch.from_numpy(g['j2d_pix']).clone())
            d['pose_out_t%d' % nt], d['tran_out_t%d' % nt] = po.reshape(T, 24, 3, 3).numpy(), to.numpy()
            print('smplify spread', name, 'threads', nt, 'moved', float((po.reshape(T, 24, 3, 3) - pose).abs().max()),
                  'vs committed golden', float((po.reshape(T, 24, 3, 3) - torch.from_numpy(g['pose_out'])).abs().max()))
        torch.set_num_threads(8)
        # ... and under float32-level noise on the start point (the size of one rounding of the inputs): 3 seeded perturbations
        for k in range(3):
            gen = torch.Generator().manual_seed(100 + k)
            tr = tran * (1 + 2e-7 * torch.randn(tran.shape, generator=gen))
            sm = ts.TemporalSMPLify(cam_k=torch.from_numpy(g['cam_k']), imu_ori=torch.from_numpy(g['imu_ori']), step_size=1e-3, num_iters=1,
                                    use_lbfgs=True, batch_size=T, max_iter=max_iter)
            po, to, rl = sm(pose.reshape(T, -1).detach(), tr.detach(), torch.from_numpy(g['j2d_pix']).clone())
            d['pose_out_eps%d' % k], d['tran_out_eps%d' % k] = po.reshape(T, 24, 3, 3).numpy(), to.numpy()
            print('smplify spread', name, 'start point perturbed by 2e-7 (relative), seed', k, 'vs committed golden',
                  float((po.reshape(T, 24, 3, 3) - torch.from_numpy(g['pose_out'])).abs().max()), float((to - torch.from_numpy(g['tran_out'])).abs().max()))
        np.savez_compressed(os.path.join(out, 'smplify_spread_%s.npz' % name), **d)


def gen_metrics(ref, out):
    """evaluate.py:120-133 (cal_mpjpe) on seeded random poses — the step right after the hot path."""
    import evaluate
    g = torch.Generator().manual_seed(41)
    T = 9
    gt = synthetic._random_rotations(T * 24, g).view(T, 24, 3, 3)
    # prediction = ground truth perturbed by small random rotations (+ one frame with a large error)
    noise = art_noise(T * 24, g, 0.15).view(T, 24, 3, 3)
    pose = gt @ noise
    pose[3] = synthetic._random_rotations(24, g)
    with torch.no_grad():
        r3 = evaluate.cal_mpjpe(pose, gt, cal_pampjpe=True)
        r2 = evaluate.cal_mpjpe(pose, gt)
    np.savez_compressed(os.path.join(out, 'metrics.npz'), pose=pose.numpy(), gt_pose=gt.numpy(), with_pa=r3.numpy(), without_pa=r2.numpy(),
                        j_regressor=evaluate.J_regressor.numpy())
    print('metrics', r3)


def gen_pipeline(ref, out):
    """SURVEY 8(f) rows 2 / 3.  (a) IMU synthesis exactly as preprocess.py:290-302 runs it (reference `_syn_acc`,
    `body_model.forward_kinematics`, `vi_mask`, `ji_mask`); (b) the dataset-row -> network-input transforms of
    evaluate.py:38-52 and 68-73, transcribed line by line on top of the reference's `articulate.math` (the statements live
    inside `evaluate_aist_ours`, which needs the AIST++ files)."""
    art = ref['art']
    import preprocess as pp
    g = torch.Generator().manual_seed(23)
    d = {}
    T = 40
    aa = (torch.randn(1, 24, 3, generator=g) * 0.3 + torch.cumsum(torch.randn(T, 24, 3, generator=g) * 0.03, dim=0))
    tran = torch.cumsum(torch.randn(T, 3, generator=g) * 0.02, dim=0)
    shape = torch.randn(10, generator=g)
    p = art.math.axis_angle_to_rotation_matrix(aa).view(-1, 24, 3, 3)
    d['imu_pose'], d['imu_tran'], d['imu_shape'] = p, tran, shape
    for tag, sh in (('mean', None), ('shaped', shape)):
        grot, joint, vert = pp.body_model.forward_kinematics(p, sh, tran, calc_mesh=True)
        d['imu_%s_vimu' % tag] = vert[:, pp.vi_mask]
        d['imu_%s_acc' % tag] = pp._syn_acc(vert[:, pp.vi_mask])                  # preprocess.py:300
        d['imu_%s_acc4' % tag] = pp._syn_acc(vert[:, pp.vi_mask], smooth_n=4)
        d['imu_%s_ori' % tag] = grot[:, pp.ji_mask]                               # preprocess.py:301
        d['imu_%s_joint' % tag] = joint
    d['imu_short_acc'] = pp._syn_acc(d['imu_mean_vimu'][:5])                      # shorter than 2 * smooth_n + 1
    # (b) two sequences, three rows (sequence 0 seen by two cameras, sequence 1 by one), ragged lengths
    lens = [17, 9]
    ori_w = [synthetic._random_rotations(L * 6, g).view(L, 6, 3, 3) for L in lens]
    acc_w = [torch.randn(L, 6, 3, generator=g) * 3 for L in lens]
    tran_w = [torch.randn(L, 3, generator=g) for L in lens]
    pose_w = [torch.randn(L, 24, 3, generator=g) * 0.4 for L in lens]
    rows = [(0, 0), (0, 1), (1, 0)]
    cam_T, cam_K, j2d = [], [], []
    for (i, j) in rows:
        Tcw = torch.eye(4)
        Tcw[:3, :3] = synthetic._random_rotations(1, g)[0]
        Tcw[:3, 3] = torch.randn(3, generator=g) * 2
        K = torch.tensor([[1000. + 50 * j, 0., 960. + 7 * i], [0., 990. + 30 * j, 540. - 5 * j], [0., 0., 1.]])
        kp = torch.rand(lens[i], 33, 3, generator=g)
        cam_T.append(Tcw); cam_K.append(K); j2d.append(kp)
    for r, (i, j) in enumerate(rows):
        Tcw, K = cam_T[r], cam_K[r]
        oric = Tcw[:3, :3].matmul(ori_w[i])                                                                   # evaluate.py:43
        accc = Tcw.matmul(art.math.append_zero(acc_w[i]).unsqueeze(-1)).squeeze(-1)[..., :3]                  # :44
        j2dc = torch.zeros(len(oric), 33, 3)                                                                  # :45-49
        j2dc[..., :2] = j2d[r][..., :2]
        j2dc[..., 0] = j2dc[..., 0] * 1920
        j2dc[..., 1] = j2dc[..., 1] * 1080
        j2dc[..., -1] = j2d[r][..., -1]
        pose = art.math.axis_angle_to_rotation_matrix(pose_w[i]).view(-1, 24, 3, 3)                           # :50-52
        pose[:, 0] = Tcw[:3, :3].matmul(pose[:, 0])
        tran = Tcw.matmul(art.math.append_one(tran_w[i]).unsqueeze(-1)).squeeze(-1)[..., :3]
        j2dn = K.inverse().matmul(art.math.append_one(j2dc[..., :2]).unsqueeze(-1)).squeeze(-1)               # :70-72
        j2dn[..., -1] = j2dc[..., -1]
        grav = Tcw[:3, :3].mm(torch.tensor([0, -1, 0.]).view(3, 1)).view(3)                                   # :73
        d['row%d_oric' % r], d['row%d_accc' % r], d['row%d_j2dc' % r] = oric, accc, j2dn
        d['row%d_pose' % r], d['row%d_tran' % r], d['row%d_gravity' % r] = pose, tran, grav
        d['row%d_cam_T' % r], d['row%d_cam_K' % r], d['row%d_j2d' % r] = Tcw, K, j2d[r]
    for i in range(2):
        d['seq%d_ori' % i], d['seq%d_acc' % i], d['seq%d_tran' % i], d['seq%d_pose_aa' % i] = ori_w[i], acc_w[i], tran_w[i], pose_w[i]
    d['rows'] = torch.tensor(rows)
    np.savez_compressed(os.path.join(out, 'pipeline.npz'), **{k: v.numpy() for k, v in d.items()})


def gen_live(ref, out):
    """the UDP text frame exactly as live_detector.py:57-61 builds it and the Unity message exactly as
    live_server.py:55-59 builds it (statements transcribed; they sit inside socket loops)."""
    import json
    art = ref['art']
    g = torch.Generator().manual_seed(5)
    frames = []
    for _ in range(3):
        uv = torch.cat((torch.randn(33, 2, generator=g) * 0.2, torch.rand(33, 1, generator=g)), dim=1)
        ori = synthetic._random_rotations(6, g)
        acc = torch.randn(6, 3, generator=g) * 4
        RCM = synthetic._random_rotations(1, g)[0]
        uvn, orin, accn, RCMn = uv.numpy(), ori.numpy(), acc.numpy(), RCM.numpy()
        data = (','.join([str(i) for i in uvn.reshape(-1)]) + '#' + ','.join(
            [str(i) for i in orin.reshape(-1)]) + '#' + ','.join([str(i) for i in accn.reshape(-1)]) + '#' + ','.join(
            [str(i) for i in RCMn.reshape(-1)])).encode()                                       # live_detector.py:57-60
        # live_server.py:42-45
        uv_s, ori_s, acc_s, rcm_s = data.decode().split('#')
        conv = lambda x: np.asarray([float(i) for i in x.split(',')])
        parsed = [torch.from_numpy(conv(x)).float() for x in (uv_s, ori_s, acc_s, rcm_s)]
        pose = synthetic._random_rotations(24, g)
        tran = torch.randn(3, generator=g) * torch.tensor([1e-3, 1.0, 1e4])
        aa = art.math.rotation_matrix_to_axis_angle(pose).view(-1)                              # live_server.py:55
        unity = ','.join(['%g' % v for v in aa]) + '#' + ','.join(['%g' % v for v in tran]) + '$'   # :57-58
        frames.append({'datagram': data.decode(), 'uv': parsed[0].tolist(), 'ori': parsed[1].tolist(), 'acc': parsed[2].tolist(),
                       'rcm': parsed[3].tolist(), 'pose_aa': [float(v) for v in aa], 'tran': [float(v) for v in tran], 'unity': unity})
    json.dump(frames, open(os.path.join(out, 'live.json'), 'w'))


def art_noise(n, g, scale):
    aa = torch.randn(n, 3, generator=g) * scale
    ang = aa.norm(dim=1, keepdim=True)
    ax = aa / ang
    K = torch.zeros(n, 3, 3)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 2], ax[:, 1], ax[:, 2], -ax[:, 0], -ax[:, 1], ax[:, 0]
    return torch.eye(3) + torch.sin(ang).unsqueeze(-1) * K + (1 - torch.cos(ang)).unsqueeze(-1) * (K @ K)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = import_reference()
    which = sys.argv[1:] or ['math', 'kinematics', 'online', 'smplify', 'metrics', 'pipeline', 'live']
    with torch.no_grad():
        if 'math' in which:
            gen_math(ref, HERE)
        if 'kinematics' in which:
            gen_kinematics(ref, HERE)
        if 'online' in which:
            gen_online(ref, HERE)
        if 'pipeline' in which:
            gen_pipeline(ref, HERE)
        if 'live' in which:
            gen_live(ref, HERE)
    if 'smplify' in which:
        gen_smplify(ref, HERE)
    if 'smplify_spread' in which or 'smplify' in which:
        gen_smplify_spread(ref, HERE)
    if 'metrics' in which:
        gen_metrics(ref, HERE)
    print('done')


if __name__ == '__main__':
    main()
