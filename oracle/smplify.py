"""Oracle (test infrastructure): sequence-level SMPLify refinement.  See ``oracle/__init__.py``.

Restates ``net/smplify/{run,temporal_smplify,losses,prior}.py``: one L-BFGS problem per sequence over the axis-angle
pose [T,72] and translation [T,3]; the objective is written out with torch ops and differentiated by autograd exactly
like the reference; the optimiser is ``torch.optim.LBFGS`` (third-party, the same class the reference calls at
``temporal_smplify.py:151``).
"""
import pickle

import numpy as np
import torch

from . import rotations as rot
from . import kinematics as kin
from .fusion import MP_MASK

JI_MASK = [18, 19, 4, 5, 15, 0]                         # config.py:101
IGNORED = [1, 2, 3, 4, 5, 6, 7, 8, 9, 31, 32]           # temporal_smplify.py:92
IGNORED_HEAD = [31, 32]                                 # temporal_smplify.py:94


class GmmPrior:
    """prior.py:83-179 (``MaxMixturePrior`` with ``use_merged=True``): min_m(0.5 d^T P_m d - log nll_w_m)."""

    def __init__(self, gmm_file, dtype=torch.float32):
        with open(gmm_file, 'rb') as f:
            g = pickle.load(f, encoding='latin1')
        np_dt = np.float32 if dtype == torch.float32 else np.float64
        means = g['means'].astype(np_dt)
        covs = g['covars'].astype(np_dt)
        self.means = torch.tensor(means, dtype=dtype)
        self.precisions = torch.tensor(np.stack([np.linalg.inv(c) for c in covs]).astype(np_dt), dtype=dtype)   # :129-130
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in g['covars']])                                      # :136-137
        const = (2 * np.pi) ** (69 / 2.)
        self.nll_weights = torch.tensor(np.asarray(g['weights'] / (const * (sqrdets / sqrdets.min()))), dtype=dtype).unsqueeze(0)

    def __call__(self, pose):
        d = pose.unsqueeze(1) - self.means                                            # :165
        pd = torch.einsum('mij,bmj->bmi', [self.precisions, d])
        quad = (pd * d).sum(dim=-1)
        ll = 0.5 * quad - torch.log(self.nll_weights)
        return ll.min(dim=1).values


def gmof(x, sigma):
    """losses.py:6-12."""
    return (sigma ** 2 * x ** 2) / (sigma ** 2 + x ** 2)


def fitting_loss(body_pose, points, joints_2d, conf, prior, cam_k, ref_points, imu_ori, ori, output='sum'):
    """losses.py:23-91 with its default weights."""
    ref = ref_points[:, 1:] - ref_points[:, :1]
    rel = points[:, 1:] - points[:, :1]
    body_3d = ((rel - ref) ** 2).sum(dim=-1)                                                        # weight 1
    proj = points / points[..., 2:]
    proj = cam_k.matmul(proj.unsqueeze(-1)).squeeze(-1)[..., :2]
    T = body_pose.shape[0]
    imu = (0.5 ** 2) * ((rot.matrix_to_axis_angle_cv2(imu_ori).reshape(T, -1) -
                         rot.matrix_to_axis_angle_cv2(ori).reshape(T, -1)) ** 2).sum(dim=-1)         # zero gradient (cv2, detached)
    reproj = (conf ** 2) * gmof(proj - joints_2d, 100).sum(dim=-1)
    pose_axis = body_pose.reshape(T, -1)[:, 3:]
    pose_prior = (0.1 ** 2) * prior(pose_axis)
    sign = torch.tensor([1., -1., -1., -1.], dtype=body_pose.dtype)
    angle = (15.2 ** 2) * (torch.exp(pose_axis[:, [52, 55, 9, 12]] * sign) ** 2).sum(dim=-1)
    # NB losses.py:63: ``imu_ori_loss`` is already [T]; its ``.sum(dim=-1)`` is a scalar that broadcasts onto every frame, so the
    # (zero-gradient) IMU term enters the summed loss T times.  Reproduced as is.
    total = reproj.sum(dim=-1) + pose_prior + angle + body_3d.sum(dim=-1) + imu.sum(dim=-1)
    cd = conf[1:]
    s2 = (cd ** 2) * (proj[1:] - proj[:-1]).abs().sum(dim=-1)
    s2 = (0.01 ** 2) * torch.cat([torch.zeros(1, s2.shape[1], dtype=s2.dtype), s2]).sum(dim=-1)
    s3 = (cd ** 2) * (points[1:] - points[:-1]).abs().sum(dim=-1)
    s3 = (1.0 ** 2) * torch.cat([torch.zeros(1, s3.shape[1], dtype=s3.dtype), s3]).sum(dim=-1)
    total = total + s2 + s3
    return total.sum() if output == 'sum' else reproj


class SmplifyOracle:
    """temporal_smplify.py:61-220."""

    def __init__(self, body, gmm_file, cam_k, imu_ori, step_size=1.0, num_iters=1, max_iter=20, use_head=False):
        self.body = body
        self.prior = GmmPrior(gmm_file)
        self.cam_k = cam_k.detach().clone()
        self.imu_ori = imu_ori.detach().clone()
        self.step_size, self.num_iters, self.max_iter = step_size, num_iters, max_iter
        self.ignored = IGNORED_HEAD if use_head else IGNORED

    def points(self, pose_mat, tran):
        gp, joint, vert = self.body.forward_kinematics(pose_mat, tran=tran, calc_mesh=True)
        return gp, kin.mediapipe_points(vert, joint, MP_MASK)

    def loss(self, body_pose, tran, joints_2d, conf, ref_points, output='sum', rodrigues=rot.batch_rodrigues):
        T = body_pose.shape[0]
        R = rodrigues(body_pose.reshape(-1, 3)).view(T, 24, 3, 3)
        gp, pts = self.points(R, tran)
        return fitting_loss(body_pose, pts, joints_2d, conf, self.prior, self.cam_k, ref_points, self.imu_ori,
                            gp[:, [JI_MASK]], output)

    def fitting_loss(self, pose_mat, tran, keypoints_2d):
        """get_fitting_loss, temporal_smplify.py:198-220 (zeroes the ignored confidences IN PLACE)."""
        T = pose_mat.shape[0]
        conf = keypoints_2d[:, :, -1]
        conf[:, self.ignored] = 0.
        body_pose = rot.matrix_to_axis_angle_cv2(pose_mat).reshape(T, -1)
        with torch.no_grad():
            gp, pts = self.points(pose_mat.reshape(T, 24, 3, 3), tran)
            return fitting_loss(body_pose, pts, keypoints_2d[:, :, :2], conf, self.prior, self.cam_k, pts, self.imu_ori,
                                gp[:, [JI_MASK]], 'reprojection')

    def __call__(self, init_pose, init_tran, keypoints_2d):
        """temporal_smplify.py:97-196."""
        T = init_pose.shape[0]
        joints_2d = keypoints_2d[:, :, :2]
        conf = keypoints_2d[:, :, -1]
        body_pose = rot.matrix_to_axis_angle_cv2(init_pose).reshape(T, -1).detach().clone()
        tran = init_tran.detach().clone()
        with torch.no_grad():
            _, ref = self.points(init_pose.reshape(T, 24, 3, 3), init_tran)
        ref = ref.detach().clone()
        tran.requires_grad = True
        body_pose.requires_grad = True
        conf[:, self.ignored] = 0.
        opt = torch.optim.LBFGS([body_pose, tran], max_iter=self.max_iter, lr=self.step_size, line_search_fn='strong_wolfe')
        self.evals = 0
        for _ in range(self.num_iters):
            def closure():
                opt.zero_grad()
                loss = self.loss(body_pose, tran, joints_2d, conf, ref)
                loss.backward()
                self.evals += 1
                return loss
            opt.step(closure)
        with torch.no_grad():
            reproj = self.loss(body_pose, tran, joints_2d, conf, ref, 'reprojection', rot.axis_angle_to_matrix)
        return rot.axis_angle_to_matrix(body_pose).detach(), tran.detach(), reproj


def smplify_runner(body, gmm_file, pred_pose, pred_tran, j2dc, imu_ori, batch_size, cam_k, lr=1.0, opt_steps=1,
                   loss_threshold=20000, use_head=False, max_iter=20):
    """run.py:6-35."""
    sm = SmplifyOracle(body, gmm_file, cam_k, imu_ori, step_size=lr, num_iters=opt_steps, max_iter=max_iter, use_head=use_head)
    pred_pose = pred_pose.reshape(batch_size, -1)
    pred_tran = pred_tran.reshape(-1, 3)
    j2dc = j2dc.reshape(-1, 33, 3)
    before = sm.fitting_loss(pred_pose.detach(), pred_tran, j2dc).mean(dim=-1)
    if before[0].sum().item() > loss_threshold:
        return pred_pose.reshape(-1, 24, 3, 3), pred_tran.reshape(-1, 3), None
    pose, tran, after = sm(pred_pose.detach(), pred_tran.detach(), j2dc)
    return pose.reshape(-1, 24, 3, 3), tran.reshape(-1, 3), after.mean(dim=-1) < before
