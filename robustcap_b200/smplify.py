r"""``smplify_runner`` / ``TemporalSMPLify`` (reference ``net/smplify/{run,temporal_smplify,losses,prior}.py``) on the B200 library.

Same call signatures and return values as the reference.  The objective of the per-sequence L-BFGS problem and its
gradient are evaluated by CUDA kernels (``csrc/smplify.cu``: analytic derivative through Rodrigues, the kinematic chain and
the skinning of the 21 vertices the 33 MediaPipe points read) instead of an autograd graph over a 6890-vertex mesh, and the optimiser —
``torch.optim.LBFGS(...).step`` with the strong-Wolfe line search, which the reference calls at temporal_smplify.py:151-166 —
runs on the device too (``rc_smplify_run``: one thread block per sequence, no host round trip per iteration;
``TemporalSMPLify.optimizer = 'torch'`` keeps the third-party class driving the native closure for A/B checks).
``smplify_runner`` gains an optional ``max_iter=20`` keyword; ``smplify_runner_batch`` refines many equal-length
sequences in one launch.
"""
import ctypes
import os
import pickle
import sys

import numpy as np
import torch

from . import _lib
from . import math as M
from .constants import SMPLIFY_IGNORED_KP, SMPLIFY_IGNORED_KP_HEAD
from .model import ParametricModel

__all__ = ['smplify_runner', 'smplify_runner_batch', 'TemporalSMPLify', 'MaxMixturePrior']


class MaxMixturePrior:
    r"""The constants of the GMM pose prior exactly as prior.py:83-157 derives them (float32 inverse of the float32
    covariances, ``nll_weights = w / ((2 pi)^34.5 * sqrtdet / min sqrtdet)``)."""

    def __init__(self, prior_folder='data/dataset_work', num_gaussians=8):
        fn = os.path.join(prior_folder, 'gmm_{:02d}.pkl'.format(num_gaussians))
        if not os.path.exists(fn):
            print('The path to the mixture prior "{}"'.format(fn) + ' does not exist, exiting!')
            sys.exit(-1)                                             # prior.py:105-108
        with open(fn, 'rb') as f:
            gmm = pickle.load(f, encoding='latin1')
        covs = gmm['covars'].astype(np.float32)
        self.means = torch.tensor(gmm['means'].astype(np.float32))
        self.precisions = torch.tensor(np.stack([np.linalg.inv(c) for c in covs]).astype(np.float32))
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in gmm['covars']])
        const = (2 * np.pi) ** (69 / 2.)
        self.nll_weights = torch.tensor(np.asarray(gmm['weights'] / (const * (sqrdets / sqrdets.min()))), dtype=torch.float32).unsqueeze(0)
        self.log_nll_weights = torch.log(self.nll_weights).reshape(-1).contiguous()


class TemporalSMPLify:
    r"""temporal_smplify.py:61-220."""
    optimizer = 'native'       # 'native': device-resident L-BFGS (rc_smplify_run); 'torch': torch.optim.LBFGS around the native closure
    body_model = None          # shared ParametricModel (module-level global in the reference, temporal_smplify.py:21)
    smpl_file = 'models/SMPL_male.pkl'

    def __init__(self, cam_k, imu_ori, step_size=1.0, num_iters=1, use_lbfgs=True, device=None, batch_size=1, max_iter=20,
                 shape=None, use_head=False, body_model=None):
        assert use_lbfgs, 'the Adam branch of the reference (temporal_smplify.py:167-180) is broken upstream and not provided'
        self.device = _lib.require_cuda()
        self.step_size, self.max_iter, self.num_iters, self.batch_size = step_size, max_iter, num_iters, batch_size
        self.cam_k = cam_k.detach().clone().to(self.device, torch.float32).contiguous()
        self.imu_ori = imu_ori.detach().clone().to(self.device, torch.float32)
        self.pose_prior = MaxMixturePrior(prior_folder='data/dataset_work', num_gaussians=8)
        self.ign_mp_joints = list(SMPLIFY_IGNORED_KP_HEAD if use_head else SMPLIFY_IGNORED_KP)
        if body_model is None:
            if TemporalSMPLify.body_model is None:
                TemporalSMPLify.body_model = ParametricModel(TemporalSMPLify.smpl_file)
            body_model = TemporalSMPLify.body_model
        # temporal_smplify.py:84-86, 158-159: an optional [10] shape vector changes the rest pose the closure skins
        self.body = body_model if shape is None else body_model.shaped(torch.as_tensor(shape))
        lib = _lib.load()
        h = _lib.vp()
        _lib.check(lib.rc_smplify_create(ctypes.byref(h), self.body._native(), _lib.hptr(self.pose_prior.means.contiguous()),
                                         _lib.hptr(self.pose_prior.precisions.contiguous()),
                                         _lib.hptr(self.pose_prior.log_nll_weights), batch_size))
        self._h = h
        self.imu_aa = M.rotation_matrix_to_axis_angle(self.imu_ori).reshape(batch_size, 18).contiguous()
        self.evals = 0

    def __del__(self):
        try:
            if getattr(self, '_h', None) is not None:
                _lib.load().rc_smplify_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _native(self, pose, tran, j2d, conf, ref3d, mode, want_grad, want_reproj=False):
        lib = _lib.load()
        T = self.batch_size
        loss = torch.empty(1, device=self.device) if want_grad else None
        g_pose = torch.empty(T, 72, device=self.device) if want_grad else None
        g_tran = torch.empty(T, 3, device=self.device) if want_grad else None
        reproj = torch.empty(T, 33, device=self.device) if want_reproj else None
        _lib.check(lib.rc_smplify_loss_grad(self._h, _lib.dptr(pose), _lib.dptr(tran), _lib.dptr(j2d), _lib.dptr(conf),
                                            _lib.dptr(self.cam_k), _lib.dptr(ref3d), _lib.dptr(self.imu_aa), mode,
                                            _lib.dptr(loss), _lib.dptr(g_pose), _lib.dptr(g_tran), _lib.dptr(reproj), _lib.stream()))
        return loss, g_pose, g_tran, reproj

    def _prep_keypoints(self, keypoints_2d):
        joints_conf = keypoints_2d[:, :, -1]
        joints_conf[:, self.ign_mp_joints] = 0.            # in place on the caller's tensor, like the reference (:148, :204)
        j2d = keypoints_2d[:, :, :2].to(self.device, torch.float32).contiguous()
        conf = joints_conf.to(self.device, torch.float32).contiguous()
        return j2d, conf

    def __call__(self, init_pose, init_tran, keypoints_2d):
        r"""temporal_smplify.py:97-196 -> (pose [T*24,3,3], tran [T,3], reprojection_loss [T,33])."""
        T = self.batch_size
        dev = self.device
        init_pose = init_pose.detach().to(dev, torch.float32).reshape(T, 24, 3, 3)
        init_tran = init_tran.detach().to(dev, torch.float32).reshape(T, 3)
        body_pose = M.rotation_matrix_to_axis_angle(init_pose).reshape(T, 72).detach().clone().contiguous()
        global_tran = init_tran.clone().contiguous()
        _, ref3d = self.body.keypoints33(init_pose, init_tran)
        ref3d = ref3d.contiguous()
        j2d, conf = self._prep_keypoints(keypoints_2d)
        if self.optimizer == 'native' and self.num_iters == 1:
            bp, gt, stats = self.optimise(body_pose, global_tran, j2d, conf, ref3d, self.imu_aa)
            self.evals += int(stats[0, 2].item())
            self.last_stats = stats
            with torch.no_grad():
                _, _, _, reproj = self._native(bp, gt, j2d, conf, ref3d, 1, False, True)
                pose = M.axis_angle_to_rotation_matrix(bp)
            return pose, gt, reproj
        body_pose.requires_grad = True
        global_tran.requires_grad = True
        opt = torch.optim.LBFGS([body_pose, global_tran], max_iter=self.max_iter, lr=self.step_size, line_search_fn='strong_wolfe')
        for _ in range(self.num_iters):
            def closure():
                opt.zero_grad()
                loss, gp, gt, _ = self._native(body_pose.detach(), global_tran.detach(), j2d, conf, ref3d, 0, True)
                body_pose.grad = gp
                global_tran.grad = gt
                self.evals += 1
                return loss[0]
            opt.step(closure)
        with torch.no_grad():
            bp, gt = body_pose.detach().contiguous(), global_tran.detach().contiguous()
            _, _, _, reproj = self._native(bp, gt, j2d, conf, ref3d, 1, False, True)
            pose = M.axis_angle_to_rotation_matrix(bp)
        return pose, gt, reproj

    def optimise(self, aa, tran, j2d, conf, ref3d, imu_aa, cam_k=None):
        r"""``torch.optim.LBFGS(max_iter, lr, strong_wolfe).step`` on the device for ``S`` sequences at once: ``aa [S*T,72]`` (or
        ``[T,72]``), ``tran [S*T,3]``, ``j2d [S*T,33,2]``, ``conf [S*T,33]``, ``ref3d [S*T,33,3]``, ``imu_aa [S*T,18]``; ``cam_k`` one
        matrix (default: the instance's) or ``[S,3,3]``.  Returns optimised ``aa``, ``tran`` and ``stats [S,4]`` =
        (first loss, final loss, closure evaluations, iterations), all on the device; no host synchronisation."""
        lib = _lib.load()
        T = self.batch_size
        S = aa.numel() // (T * 72)
        dev = self.device
        c = lambda x: x.detach().to(dev, torch.float32).contiguous()
        aa, tran, j2d, conf, ref3d, imu_aa = c(aa), c(tran), c(j2d), c(conf), c(ref3d), c(imu_aa)
        cam = self.cam_k if cam_k is None else c(cam_k)
        per_seq = int(cam.numel() == 9 * S and S > 1)
        out_aa, out_tran = torch.empty_like(aa), torch.empty_like(tran)
        stats = torch.empty(S, 4, device=dev)
        _lib.check(lib.rc_smplify_run(self._h, S, _lib.dptr(aa), _lib.dptr(tran), _lib.dptr(j2d), _lib.dptr(conf), _lib.dptr(cam), per_seq,
                                      _lib.dptr(ref3d), _lib.dptr(imu_aa), int(self.max_iter), float(self.step_size),
                                      _lib.dptr(out_aa), _lib.dptr(out_tran), _lib.dptr(stats), _lib.stream()))
        self._keep = (aa, tran, j2d, conf, ref3d, imu_aa, cam)
        return out_aa, out_tran, stats

    def get_fitting_loss(self, pose, tran, keypoints_2d):
        r"""temporal_smplify.py:198-220 -> reprojection loss [T,33] of the given rotation matrices."""
        T = self.batch_size
        pose = pose.detach().to(self.device, torch.float32).reshape(T, 216).contiguous()
        tran = tran.detach().to(self.device, torch.float32).reshape(T, 3).contiguous()
        j2d, conf = self._prep_keypoints(keypoints_2d)
        dummy = torch.zeros(T, 33, 3, device=self.device)
        return self._native(pose, tran, j2d, conf, dummy, 2, False, True)[3]


def smplify_runner_batch(pred_pose, pred_tran, j2dc, imu_ori, cam_k, lr=1.0, max_iter=20, use_head=False, body_model=None):
    r"""``smplify_runner`` (run.py:6-35 with ``loss_threshold = inf``) for ``S`` equal-length sequences at once — the batched offline
    evaluation of BASELINE configs[3]: ``pred_pose [S,T,24,3,3]``, ``pred_tran [S,T,3]``, ``j2dc [S,T,33,3]`` pixel key points with
    confidence, ``imu_ori [S,T,6,3,3]``, ``cam_k [3,3]`` or ``[S,3,3]``.  Everything stays on the device; returns
    ``pose [S,T,24,3,3]``, ``tran [S,T,3]``, ``stats [S,4]`` (first loss, final loss, closure evaluations, iterations)."""
    S, T = pred_pose.shape[0], pred_pose.shape[1]
    sm = TemporalSMPLify(cam_k=cam_k.reshape(-1, 3, 3)[0], imu_ori=imu_ori.reshape(S * T, 6, 3, 3)[:T], step_size=lr, batch_size=T,
                         max_iter=max_iter, use_head=use_head, body_model=body_model)
    dev = sm.device
    pose = pred_pose.detach().to(dev, torch.float32).reshape(S * T, 24, 3, 3)
    tran = pred_tran.detach().to(dev, torch.float32).reshape(S * T, 3)
    aa = M.rotation_matrix_to_axis_angle(pose).reshape(S * T, 72)
    _, ref3d = sm.body.keypoints33(pose, tran)
    kp = j2dc.detach().to(dev, torch.float32).reshape(S * T, 33, 3).clone()
    kp[:, sm.ign_mp_joints, 2] = 0.
    imu_aa = M.rotation_matrix_to_axis_angle(imu_ori.detach().to(dev, torch.float32).reshape(S * T * 6, 3, 3)).reshape(S * T, 18)
    aa2, tran2, stats = sm.optimise(aa, tran, kp[:, :, :2], kp[:, :, 2], ref3d, imu_aa, cam_k=cam_k)
    return M.axis_angle_to_rotation_matrix(aa2).reshape(S, T, 24, 3, 3), tran2.reshape(S, T, 3), stats


def smplify_runner(pred_pose, pred_tran, j2dc, imu_ori, batch_size, cam_k, lr=1.0, opt_steps=1, use_lbfgs=True,
                   loss_threshold=20000, shape=None, use_head=False, max_iter=20):
    r"""run.py:6-35 -> (pose [T,24,3,3] cpu, tran [T,3] cpu, update bool[T] or None)."""
    smplify = TemporalSMPLify(step_size=lr, batch_size=batch_size, num_iters=opt_steps, use_lbfgs=use_lbfgs, cam_k=cam_k,
                              imu_ori=imu_ori, shape=shape, use_head=use_head, max_iter=max_iter)
    pred_pose = pred_pose.reshape(batch_size, -1)
    pred_tran = pred_tran.reshape(-1, 3)
    j2dc = j2dc.reshape(-1, 33, 3)
    opt_joint_loss = smplify.get_fitting_loss(pred_pose.detach(), pred_tran, j2dc).mean(dim=-1)
    if opt_joint_loss[0].sum().cpu().item() > loss_threshold:
        return pred_pose.cpu().reshape(-1, 24, 3, 3), pred_tran.cpu().reshape(-1, 3), None
    pose, tran, new_opt_joint_loss = smplify(pred_pose.detach(), pred_tran.detach(), j2dc)
    update = (new_opt_joint_loss.mean(dim=-1) < opt_joint_loss)
    return pose.cpu().reshape(-1, 24, 3, 3), tran.cpu().reshape(-1, 3), update
