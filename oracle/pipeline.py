"""CPU oracle (TEST INFRASTRUCTURE ONLY — never imported by the product) of SURVEY.md 8(f) rows 2 / 3:

* ``pack_row``          evaluate.py:38-52, 68-73 — one (sequence, camera) dataset row -> network inputs and labels.
* ``syn_acc``           preprocess.py:22-33 (``_syn_acc``).
* ``synthesize_imu``    preprocess.py:290-302 on top of ``oracle.kinematics.BodyOracle``.

Pinned against outputs of the unmodified reference (tests/golden/pipeline.npz, tests/golden/make_golden.py::gen_pipeline).
"""
import torch

from .rotations import axis_angle_to_matrix

VI_MASK = [1961, 5424, 1176, 4662, 411, 3021]     # config.py:100
JI_MASK = [18, 19, 4, 5, 15, 0]                   # config.py:101


def pack_row(j2d, imu_acc, imu_ori, cam_T, cam_K, image_size=(1920, 1080), pose_aa=None, tran=None):
    """evaluate.py:38-52 (dataset construction) followed by :68-73 (per-sequence normalisation)."""
    R = cam_T[:3, :3]
    oric = R.matmul(imu_ori)                                                              # :43
    accc = cam_T.matmul(torch.cat((imu_acc, torch.zeros_like(imu_acc[..., :1])), -1).unsqueeze(-1)).squeeze(-1)[..., :3]   # :44
    px = torch.zeros(len(oric), 33, 3, dtype=j2d.dtype)                                   # :45-49
    px[..., 0] = j2d[..., 0] * image_size[0]
    px[..., 1] = j2d[..., 1] * image_size[1]
    px[..., 2] = j2d[..., 2]
    j2dc = cam_K.inverse().matmul(torch.cat((px[..., :2], torch.ones_like(px[..., :1])), -1).unsqueeze(-1)).squeeze(-1)    # :70-71
    j2dc[..., -1] = px[..., -1]                                                           # :72
    gravity = R.mm(torch.tensor([0, -1, 0.], dtype=R.dtype).view(3, 1)).view(3)           # :73
    out = {'j2dc': j2dc, 'accc': accc, 'oric': oric, 'gravity': gravity}
    if pose_aa is not None:                                                               # :50-52
        pose = axis_angle_to_matrix(pose_aa.reshape(-1, 3)).view(-1, 24, 3, 3)
        pose[:, 0] = R.matmul(pose[:, 0])
        out['pose'] = pose
    if tran is not None:
        out['tran'] = cam_T.matmul(torch.cat((tran, torch.ones_like(tran[..., :1])), -1).unsqueeze(-1)).squeeze(-1)[..., :3]
    return out


def syn_acc(v, smooth_n=2):
    """preprocess.py:22-33 restated with slices: central second differences at 60 fps (x 3600), zero at both ends; away from the ends
    (``smooth_n <= t < T - smooth_n``) the stencil is widened to +-smooth_n frames and divided by smooth_n^2.  Element for element the
    same float32 operations in the same order as the reference's list comprehensions."""
    T = v.shape[0]
    acc = torch.zeros_like(v)
    acc[1:T - 1] = (v[:T - 2] + v[2:] - 2 * v[1:T - 1]) * 3600
    s = smooth_n
    if s // 2 != 0 and T > 2 * s:
        acc[s:T - s] = (v[:T - 2 * s] + v[2 * s:] - 2 * v[s:T - s]) * 3600 / s ** 2
    return acc


def synthesize_imu(body, pose, tran=None, shape=None, smooth_n=2):
    """preprocess.py:290-302: (imu_acc [T,6,3], imu_ori [T,6,3,3], joint [T,24,3], vimu [T,6,3])."""
    grot, joint, vert = body.forward_kinematics(pose, shape, tran, calc_mesh=True)
    vimu = vert[:, VI_MASK]
    return syn_acc(vimu, smooth_n), grot[:, JI_MASK], joint, vimu
