"""Oracle (test infrastructure): evaluation metrics next to the hot path.  See ``oracle/__init__.py``.

Restates ``evaluate.py:120-133`` (``cal_mpjpe``) and the Procrustes alignment it calls, ``utils.py:138-203``
(``compute_similarity_transform`` / ``reconstruction_error``), SURVEY.md §8(f) row 1.
"""
import numpy as np
import torch


def similarity_align(S1, S2):
    """utils.py:138-186 for one [N,3] pair: the similarity transform (s, R, t) of S1 closest to S2, returns the aligned S1."""
    X1, X2 = S1.T, S2.T
    mu1, mu2 = X1.mean(axis=1, keepdims=True), X2.mean(axis=1, keepdims=True)
    Y1, Y2 = X1 - mu1, X2 - mu2
    var1 = np.sum(Y1 ** 2)
    K = Y1.dot(Y2.T)
    U, s, Vh = np.linalg.svd(K)
    V = Vh.T
    Z = np.eye(3)
    Z[-1, -1] *= np.sign(np.linalg.det(U.dot(V.T)))
    R = V.dot(Z.dot(U.T))
    scale = np.trace(R.dot(K)) / var1
    t = mu2 - scale * (R.dot(mu1))
    return (scale * R.dot(X1) + t).T


def reconstruction_error(S1, S2):
    """utils.py:195-203 with reduction=None: per-sample mean joint distance after Procrustes alignment."""
    out = np.zeros(S1.shape[0], dtype=S1.dtype)
    for i in range(S1.shape[0]):
        out[i] = np.sqrt(((similarity_align(S1[i], S2[i]) - S2[i]) ** 2).sum(axis=-1)).mean()
    return out


def cal_mpjpe(body, j_regressor, pose, gt_pose, cal_pampjpe=False):
    """evaluate.py:120-133: (MPJPE over the first 14 regressed joints after pelvis alignment, PVE, [PA-MPJPE])."""
    _, _, gv = body.forward_kinematics(gt_pose, calc_mesh=True)
    _, _, pv = body.forward_kinematics(pose, calc_mesh=True)
    gk = torch.matmul(j_regressor, gv)[:, :14]
    pk = torch.matmul(j_regressor, pv)[:, :14]
    pk = pk - pk[:, [0]]
    gk = gk - gk[:, [0]]
    res = [(gk - pk).norm(dim=2).mean(), (gv - pv).norm(dim=2).mean()]
    if cal_pampjpe:
        res.append(torch.tensor(reconstruction_error(pk.numpy(), gk.numpy()).mean()))
    return torch.stack([torch.as_tensor(r, dtype=torch.float32) for r in res])
