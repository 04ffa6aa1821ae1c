r"""Evaluation metrics next to the hot path: ``cal_mpjpe`` of the reference's ``evaluate.py:120-133`` on the GPU.

Both meshes are skinned, regressed to the 14 evaluation joints and reduced inside one kernel (no 6890-vertex tensor is
materialised); the Procrustes alignment of ``utils.py:138-203`` runs per frame on the device (Horn's closed form in float64).
"""
import torch

from . import _lib

__all__ = ['cal_mpjpe', 'reconstruction_error']


def cal_mpjpe(body_model, j_regressor, pose, gt_pose, cal_pampjpe=False):
    r"""-> tensor [MPJPE, PVE] or [MPJPE, PVE, PA-MPJPE] (metres), same reductions as the reference."""
    lib = _lib.load()
    dev = _lib.require_cuda()
    T = pose.shape[0]
    p = pose.detach().reshape(T, 216).to(dev, torch.float32).contiguous()
    g = gt_pose.detach().reshape(T, 216).to(dev, torch.float32).contiguous()
    jr = j_regressor.detach().to(dev, torch.float32).contiguous()
    out = torch.empty(T, 3, device=dev)
    _lib.check(lib.rc_metrics_mpjpe(body_model._native(), _lib.dptr(jr), jr.shape[0], _lib.dptr(p), _lib.dptr(g), T, int(bool(cal_pampjpe)),
                                    _lib.dptr(out), _lib.stream()))
    m = out.mean(dim=0).cpu()
    return m if cal_pampjpe else m[:2]


def reconstruction_error(S1, S2, reduction='mean'):
    r"""utils.py:195-203 kept for the drop-in ``utils`` module (host numpy; evaluate.py only reaches it through cal_mpjpe)."""
    import numpy as np
    S1, S2 = np.asarray(S1), np.asarray(S2)
    out = np.zeros(S1.shape[0], dtype=S1.dtype)
    for i in range(S1.shape[0]):
        X1, X2 = S1[i].T, S2[i].T
        mu1, mu2 = X1.mean(axis=1, keepdims=True), X2.mean(axis=1, keepdims=True)
        Y1, Y2 = X1 - mu1, X2 - mu2
        K = Y1.dot(Y2.T)
        U, s, Vh = np.linalg.svd(K)
        V = Vh.T
        Z = np.eye(3)
        Z[-1, -1] *= np.sign(np.linalg.det(U.dot(V.T)))
        R = V.dot(Z.dot(U.T))
        scale = np.trace(R.dot(K)) / np.sum(Y1 ** 2)
        t = mu2 - scale * (R.dot(mu1))
        out[i] = np.sqrt((((scale * R.dot(X1) + t).T - S2[i]) ** 2).sum(axis=-1)).mean()
    if reduction == 'mean':
        return out.mean()
    if reduction == 'sum':
        return out.sum()
    return out
