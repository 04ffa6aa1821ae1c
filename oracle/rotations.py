"""Oracle (test infrastructure): rotation-representation conversions.  See ``oracle/__init__.py``.

Restates ``articulate/math/general.py`` and ``articulate/math/angular.py`` of the reference.
All functions take/return torch CPU tensors and keep the input dtype (float32 mirrors the reference;
float64 gives a higher-precision "truth" to size rounding noise against).
"""
import math

import torch


def lerp(a, b, t):
    """general.py:15-24."""
    return a * (1 - t) + b * t


def unit(x, dim=-1):
    """general.py:27-39 (``normalize_tensor``): x / ||x||, no epsilon (0-vectors give NaN)."""
    return x / x.norm(dim=dim, keepdim=True)


def skew(v):
    """general.py:63-74: [v]x for each row of v."""
    v = v.reshape(-1, 3)
    z = torch.zeros_like(v[:, 0])
    return torch.stack((z, -v[:, 2], v[:, 1], v[:, 2], z, -v[:, 0], -v[:, 1], v[:, 0], z), dim=1).view(-1, 3, 3)


def r6d_to_matrix(r6d):
    """angular.py:249-264: Gram-Schmidt, the three unit vectors become COLUMNS, NaN -> 0."""
    x = r6d.reshape(-1, 6)
    a, b = x[:, :3], x[:, 3:]
    c0 = unit(a)
    c1 = unit(b - (c0 * b).sum(dim=1, keepdim=True) * c0)
    c2 = torch.linalg.cross(c0, c1, dim=1)
    r = torch.stack((c0, c1, c2), dim=-1)
    return torch.where(torch.isnan(r), torch.zeros_like(r), r)


def matrix_to_r6d(r):
    """angular.py:267-274: first two columns, column 0 then column 1."""
    r = r.reshape(-1, 3, 3)
    return torch.cat((r[:, :, 0], r[:, :, 1]), dim=1).clone()


def axis_angle_to_matrix(a):
    """angular.py:221-233: c*I + (1-c)*a a^T + s*[a]x with the axis' NaN/Inf zeroed."""
    a = a.reshape(-1, 3)
    angle = a.norm(dim=1, keepdim=True)
    axis = a / angle
    axis = torch.where(torch.isnan(axis) | torch.isinf(axis), torch.zeros_like(axis), axis)
    c, s = angle.cos().view(-1, 1, 1), angle.sin().view(-1, 1, 1)
    eye = torch.eye(3, dtype=a.dtype).expand(a.shape[0], 3, 3)
    return c * eye + (1 - c) * axis.unsqueeze(2) * axis.unsqueeze(1) + s * skew(axis)


def batch_rodrigues(a):
    """temporal_smplify.py:25-59: angle = ||a + 1e-8||, R = I + sin*K + (1-cos)*K@K (differentiable)."""
    a = a.reshape(-1, 3)
    angle = torch.norm(a + 1e-8, dim=1, keepdim=True)
    d = a / angle
    K = skew(d)
    s, c = torch.sin(angle).unsqueeze(1), torch.cos(angle).unsqueeze(1)
    eye = torch.eye(3, dtype=a.dtype).unsqueeze(0)
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def matrix_to_axis_angle_cv2(r):
    """angular.py:236-246: literally ``cv2.Rodrigues`` per matrix (the third-party routine the reference calls)."""
    import cv2
    import numpy as np
    m = r.detach().reshape(-1, 3, 3).cpu().numpy()
    out = np.stack([cv2.Rodrigues(x)[0] for x in m])
    return torch.from_numpy(out).to(r.dtype).squeeze(-1)


def matrix_to_axis_angle(r):
    """Restatement of what ``cv2.Rodrigues(matrix)`` computes (OpenCV ``calib3d`` Rodrigues, matrix branch;
    the pinned dependency is opencv 4.2.0.34, the container has 4.13 — same algorithm), in float64 internally
    exactly as OpenCV does:

    1. project onto SO(3): ``R <- U @ Vt`` of the SVD;
    2. ``r = (R32-R23, R13-R31, R21-R12)``, ``s = |r|/2``, ``c = (tr R - 1)/2`` clamped to [-1, 1], ``theta = acos(c)``;
    3. if ``s < 1e-5``: ``c > 0`` -> zero vector; else (theta ~ pi) take the axis from the diagonal
       (``sqrt(max((Rii+1)/2, 0))``) with signs fixed from ``R12``, ``R13``, ``R23`` and scale by ``theta/|axis|``;
       otherwise ``r * theta / (2 s)``.

    Checked against ``matrix_to_axis_angle_cv2`` in ``tests/test_oracle_golden.py``.
    """
    dt = r.dtype
    R = r.detach().reshape(-1, 3, 3).to(torch.float64)
    U, _, Vt = torch.linalg.svd(R)
    R = U @ Vt
    rx = R[:, 2, 1] - R[:, 1, 2]
    ry = R[:, 0, 2] - R[:, 2, 0]
    rz = R[:, 1, 0] - R[:, 0, 1]
    s = torch.sqrt((rx * rx + ry * ry + rz * rz) * 0.25)
    c = ((R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2] - 1) * 0.5).clamp(-1, 1)
    theta = torch.acos(c)
    out = torch.zeros(R.shape[0], 3, dtype=torch.float64)
    small = s < 1e-5
    # generic branch
    vth = torch.where(small, torch.zeros_like(s), 1.0 / (2 * torch.where(small, torch.ones_like(s), s))) * theta
    out = torch.stack((rx, ry, rz), dim=1) * vth.unsqueeze(1)
    # theta ~ pi branch
    near_pi = small & (c <= 0)
    if near_pi.any():
        t = torch.sqrt(torch.clamp((R[:, 0, 0] + 1) * 0.5, min=0))
        ay = torch.sqrt(torch.clamp((R[:, 1, 1] + 1) * 0.5, min=0)) * torch.where(R[:, 0, 1] < 0, -1.0, 1.0)
        az = torch.sqrt(torch.clamp((R[:, 2, 2] + 1) * 0.5, min=0)) * torch.where(R[:, 0, 2] < 0, -1.0, 1.0)
        flip = ((t.abs() < ay.abs()) & (t.abs() < az.abs())) & ((R[:, 1, 2] > 0) != (ay * az > 0))
        az = torch.where(flip, -az, az)
        n = torch.sqrt(t * t + ay * ay + az * az)
        alt = torch.stack((t, ay, az), dim=1) * (theta / n).unsqueeze(1)
        out = torch.where(near_pi.unsqueeze(1), alt, out)
    out = torch.where((small & (c > 0)).unsqueeze(1), torch.zeros_like(out), out)
    return out.to(dt)


def quaternion_to_matrix(q):
    """angular.py:306-318 (wxyz, normalised first)."""
    q = unit(q.reshape(-1, 4))
    a, b, c, d = q[:, 0:1], q[:, 1:2], q[:, 2:3], q[:, 3:4]
    r = torch.cat((-2 * c * c - 2 * d * d + 1, 2 * b * c - 2 * a * d, 2 * a * c + 2 * b * d,
                   2 * b * c + 2 * a * d, -2 * b * b - 2 * d * d + 1, 2 * c * d - 2 * a * b,
                   2 * b * d - 2 * a * c, 2 * a * b + 2 * c * d, -2 * b * b - 2 * c * c + 1), dim=1)
    return r.view(-1, 3, 3)


def quaternion_to_axis_angle(q):
    """angular.py:277-290: may return angles in (pi, 2pi); NaN -> 0."""
    q = unit(q.reshape(-1, 4))
    th = q[:, 0].clamp(-1, 1).acos()
    a = q[:, 1:] / th.sin().view(-1, 1) * 2 * th.view(-1, 1)
    return torch.where(torch.isnan(a), torch.zeros_like(a), a)


def axis_angle_to_quaternion(a):
    """angular.py:293-303."""
    a = a.reshape(-1, 3)
    ang = a.norm(dim=1, keepdim=True)
    ax = a / ang
    ax = torch.where(torch.isnan(ax), torch.zeros_like(ax), ax)
    return torch.cat(((ang / 2).cos(), (ang / 2).sin() * ax), dim=1)


def quaternion_product(q1, q2):
    """angular.py:79-93."""
    a, b = q1.reshape(-1, 4), q2.reshape(-1, 4)
    w1, v1, w2, v2 = a[:, :1], a[:, 1:], b[:, :1], b[:, 1:]
    v = torch.linalg.cross(v1, v2, dim=1) + w1 * v2 + w2 * v1
    w = w1 * w2 - (v1 * v2).sum(dim=1, keepdim=True)
    return torch.cat((w, v), dim=1).view_as(q1)


def quaternion_inverse(q):
    """angular.py:96-106 (conjugate)."""
    out = q.clone().reshape(-1, 4)
    out[:, 1:] = -out[:, 1:]
    return out.view_as(q)


def quaternion_mean(q):
    """angular.py:68-77."""
    q = q.reshape(-1, 4)
    k = int(q.abs().mean(dim=0).argmax())
    q = q * q[:, k].sign().view(-1, 1)
    return unit(q.mean(dim=0))


def angle_between(r1, r2):
    """angular.py:128-141 for rotation matrices: |log(R1^T R2)|."""
    off = r1.reshape(-1, 3, 3).transpose(1, 2).bmm(r2.reshape(-1, 3, 3))
    return matrix_to_axis_angle_cv2(off).norm(dim=1)
