r"""``articulate.math`` (reference ``articulate/math/{general,angular,spatial}.py``) on the B200 library.

Same names, argument meaning and return shapes as the reference's ``__all__`` lists.  Every arithmetic routine on
the hot path runs in the CUDA library through the C ABI; tensors that arrive on the CPU are moved to the GPU,
processed there and the result is returned on the caller's device (there is no CPU implementation — without a
CUDA device these functions raise).  Pure tensor plumbing (``lerp``, ``append_value`` ...) stays in torch.
"""
import enum
from functools import partial

import numpy as np
import torch

from . import _lib

__all__ = ['lerp', 'normalize_tensor', 'append_value', 'append_zero', 'append_one', 'vector_cross_matrix',
           'vector_cross_matrix_np', 'block_diagonal_matrix_np',
           'RotationRepresentation', 'to_rotation_matrix', 'radian_to_degree', 'degree_to_radian', 'normalize_angle',
           'angle_difference', 'angle_between', 'svd_rotate', 'generate_random_rotation_matrix',
           'axis_angle_to_rotation_matrix', 'rotation_matrix_to_axis_angle', 'r6d_to_rotation_matrix',
           'rotation_matrix_to_r6d', 'quaternion_to_axis_angle', 'axis_angle_to_quaternion',
           'quaternion_to_rotation_matrix', 'rotation_matrix_to_euler_angle', 'euler_angle_to_rotation_matrix',
           'rotation_matrix_to_euler_angle_np', 'euler_angle_to_rotation_matrix_np', 'euler_convert_np',
           'quaternion_product', 'quaternion_inverse', 'quaternion_mean', 'generate_random_rotation_matrix_constrained',
           'transformation_matrix_np', 'adjoint_transformation_matrix_np', 'transformation_matrix',
           'decode_transformation_matrix', 'inverse_transformation_matrix', 'bone_vector_to_joint_position',
           'joint_position_to_bone_vector', 'forward_kinematics_R', 'inverse_kinematics_R', 'forward_kinematics_T',
           'inverse_kinematics_T', 'forward_kinematics', 'batch_rodrigues']


# ---------------------------------------------------------------------------------------------------------------
# native dispatch helpers
def _map(name, x, in_w, out_w, out_shape=None):
    """Run the item-wise kernel ``name`` over ``x`` viewed as [-1, in_w]."""
    lib = _lib.load()
    dev = _lib.require_cuda()
    src = x.detach().reshape(-1, in_w)
    d = src.to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty(d.shape[0], out_w, device=dev, dtype=torch.float32)
    _lib.check(getattr(lib, name)(_lib.dptr(d), _lib.dptr(out), d.shape[0], _lib.stream()))
    out = out.to(x.device)
    return out if out_shape is None else out.view(*out_shape)


def _parent_array(parent):
    p = [-1 if (v is None or v < 0) else int(v) for v in parent]
    p[0] = -1
    return torch.tensor(p, dtype=torch.int32)


def _tree(name, x, parent, elem):
    lib = _lib.load()
    dev = _lib.require_cuda()
    nj = len(parent)
    src = x.detach().reshape(x.shape[0], nj * elem)
    d = src.to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty_like(d)
    par = _parent_array(parent)
    _lib.check(getattr(lib, name)(_lib.dptr(d), _lib.dptr(out), _lib.hptr(par), nj, d.shape[0], _lib.stream()))
    return out.to(x.device)


# ---------------------------------------------------------------------------------------------------------------
# general.py
def lerp(a, b, t):
    r"""Linear interpolation (unclamped). general.py:15-24."""
    return a * (1 - t) + b * t


def normalize_tensor(x: torch.Tensor, dim=-1, return_norm=False):
    r"""Normalize a tensor in a specific dimension to unit norm. general.py:27-39."""
    norm = x.norm(dim=dim, keepdim=True)
    y = x / norm
    return y if not return_norm else (y, norm)


def append_value(x: torch.Tensor, value: float, dim=-1):
    r"""Append a value to a tensor in a specific dimension. general.py:42-56."""
    shape = list(x.shape)
    shape[dim] = 1
    return torch.cat((x, torch.full(shape, value, dtype=x.dtype, device=x.device)), dim=dim)


append_zero = partial(append_value, value=0)
append_one = partial(append_value, value=1)


def vector_cross_matrix(x: torch.Tensor):
    r"""Skew-symmetric matrix of each vector3. general.py:63-74."""
    x = x.view(-1, 3)
    z = torch.zeros(x.shape[0], device=x.device, dtype=x.dtype)
    return torch.stack((z, -x[:, 2], x[:, 1], x[:, 2], z, -x[:, 0], -x[:, 1], x[:, 0], z), dim=1).view(-1, 3, 3)


def vector_cross_matrix_np(x):
    r"""general.py:77-86."""
    return np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]], dtype=float)


def block_diagonal_matrix_np(matrix2d_list):
    r"""general.py:89-103."""
    ret = np.zeros(sum([np.array(m.shape) for m in matrix2d_list]))
    r = c = 0
    for m in matrix2d_list:
        ret[r:r + m.shape[0], c:c + m.shape[1]] = m
        r, c = r + m.shape[0], c + m.shape[1]
    return ret


# ---------------------------------------------------------------------------------------------------------------
# angular.py
class RotationRepresentation(enum.Enum):
    r"""Quaternions are wxyz, Euler angles local XYZ. angular.py:21-29."""
    AXIS_ANGLE = 0
    ROTATION_MATRIX = 1
    QUATERNION = 2
    R6D = 3
    EULER_ANGLE = 4


def to_rotation_matrix(r: torch.Tensor, rep: RotationRepresentation):
    r"""angular.py:32-51."""
    if rep == RotationRepresentation.AXIS_ANGLE:
        return axis_angle_to_rotation_matrix(r)
    if rep == RotationRepresentation.QUATERNION:
        return quaternion_to_rotation_matrix(r)
    if rep == RotationRepresentation.R6D:
        return r6d_to_rotation_matrix(r)
    if rep == RotationRepresentation.EULER_ANGLE:
        return euler_angle_to_rotation_matrix(r)
    if rep == RotationRepresentation.ROTATION_MATRIX:
        return r.view(-1, 3, 3)
    raise Exception('unknown rotation representation')


def radian_to_degree(q):
    return q * 180.0 / np.pi


def degree_to_radian(q):
    return q / 180.0 * np.pi


def normalize_angle(q):
    r"""Into [-pi, pi). angular.py:109-118."""
    mod = q % (2 * np.pi)
    mod[mod >= np.pi] -= 2 * np.pi
    return mod


def angle_difference(target, source):
    return normalize_angle(target - source)


def axis_angle_to_rotation_matrix(a: torch.Tensor):
    r"""[*, 3] -> [N, 3, 3]. angular.py:221-233."""
    return _map('rc_axis_angle_to_rotmat', a, 3, 9, (-1, 3, 3))


def rotation_matrix_to_axis_angle(r: torch.Tensor):
    r"""[*, 3, 3] -> [N, 3] with cv2.Rodrigues semantics (SO(3) projection first), detached. angular.py:236-246."""
    return _map('rc_rotmat_to_axis_angle', r, 9, 3)


def r6d_to_rotation_matrix(r6d: torch.Tensor):
    r"""[*, 6] -> [N, 3, 3] (Gram-Schmidt, columns). angular.py:249-264."""
    return _map('rc_r6d_to_rotmat', r6d, 6, 9, (-1, 3, 3))


def rotation_matrix_to_r6d(r: torch.Tensor):
    r"""[*, 3, 3] -> [N, 6]. angular.py:267-274."""
    return _map('rc_rotmat_to_r6d', r, 9, 6)


def batch_rodrigues(rot_vecs: torch.Tensor, epsilon=1e-8):
    r"""[N, 3] -> [N, 3, 3], the SPIN variant used inside SMPLify. net/smplify/temporal_smplify.py:25-59."""
    return _map('rc_batch_rodrigues', rot_vecs, 3, 9, (-1, 3, 3))


def quaternion_to_axis_angle(q: torch.Tensor):
    r"""angular.py:277-290."""
    return _map('rc_quat_to_axis_angle', q, 4, 3)


def axis_angle_to_quaternion(a: torch.Tensor):
    r"""angular.py:293-303."""
    return _map('rc_axis_angle_to_quat', a, 3, 4)


def quaternion_to_rotation_matrix(q: torch.Tensor):
    r"""angular.py:306-318."""
    return _map('rc_quat_to_rotmat', q, 4, 9, (-1, 3, 3))


def quaternion_product(q1, q2):
    r"""angular.py:79-93."""
    lib = _lib.load()
    dev = _lib.require_cuda()
    a = q1.detach().reshape(-1, 4).to(device=dev, dtype=torch.float32).contiguous()
    b = q2.detach().reshape(-1, 4).to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty_like(a)
    _lib.check(lib.rc_quat_product(_lib.dptr(a), _lib.dptr(b), _lib.dptr(out), a.shape[0], _lib.stream()))
    return out.to(q1.device).view_as(q1)


def quaternion_inverse(q):
    r"""Conjugate. angular.py:96-106."""
    inv = q.clone().view(-1, 4)
    inv[:, 1:].neg_()
    return inv.view_as(q)


def quaternion_mean(q):
    r"""angular.py:68-77."""
    q = q.reshape(-1, 4)
    q = q * q[:, int(q.abs().mean(dim=0).argmax())].sign().view(-1, 1).expand(-1, 4)
    return normalize_tensor(q.mean(dim=0))


def angle_between(rot1: torch.Tensor, rot2: torch.Tensor, rep=RotationRepresentation.ROTATION_MATRIX):
    r"""Angle in radians between two rotations. angular.py:128-141."""
    rot1 = to_rotation_matrix(rot1, rep)
    rot2 = to_rotation_matrix(rot2, rep)
    offsets = rot1.transpose(1, 2).bmm(rot2)
    return rotation_matrix_to_axis_angle(offsets).norm(dim=1)


def svd_rotate(source_points: torch.Tensor, target_points: torch.Tensor, calc_R=True, calc_t=False, calc_s=False):
    r"""Similarity alignment (used by metrics/training only, host torch ops). angular.py:144-184."""
    zeros = torch.zeros_like(source_points[:, :1])
    sm = source_points.mean(dim=1, keepdim=True) if calc_t else zeros
    tm = target_points.mean(dim=1, keepdim=True) if calc_t else zeros
    if calc_s:
        scale = (((target_points - tm) ** 2).sum(dim=[1, 2]) / ((source_points - sm) ** 2).sum(dim=[1, 2])).sqrt()
    else:
        scale = torch.ones_like(source_points[:, 0, 0])
    if calc_R:
        u, _, vh = torch.linalg.svd((source_points - sm).transpose(1, 2).bmm(target_points - tm))
        v = vh.transpose(1, 2).clone()
        vut = v.bmm(u.transpose(1, 2))
        flip = torch.linalg.det(vut) < -0.9
        v[flip, 2] = -v[flip, 2]     # reference negates row 2 of V (angular.py:173), kept as is
        vut = v.bmm(u.transpose(1, 2))
        rotation = vut
    else:
        rotation = torch.eye(source_points.shape[2], device=source_points.device).repeat(source_points.shape[0], 1, 1)
    translation = -scale.view(-1, 1, 1) * rotation.bmm(sm.transpose(1, 2)) + tm.transpose(1, 2)
    transformed = scale.view(-1, 1, 1) * source_points.bmm(rotation.transpose(1, 2)) + translation.transpose(1, 2)
    return rotation, translation.squeeze(2), scale, transformed


def generate_random_rotation_matrix(n=1):
    r"""Uniform random rotations by quaternion rejection sampling. angular.py:187-202."""
    q = torch.zeros(n, 4)
    while True:
        nrm = q.norm(dim=1)
        mask = (nrm == 0) | (nrm > 1)
        if not mask.any():
            break
        q[mask] = torch.rand_like(q[mask]) * 2 - 1
    return quaternion_to_rotation_matrix(q / q.norm(dim=1, keepdim=True))


def generate_random_rotation_matrix_constrained(n=1, y=(-180, 180), p=(-90, 90), r=(-180, 180)):
    r"""angular.py:205-218."""
    ry = degree_to_radian(lerp(y[0], y[1], torch.rand(n)))
    rp = degree_to_radian(lerp(p[0], p[1], torch.rand(n)))
    rr = degree_to_radian(lerp(r[0], r[1], torch.rand(n)))
    return euler_angle_to_rotation_matrix(torch.stack((ry, rp, rr), dim=1), seq='YXZ')


# Euler conversions are off the hot path (dataset preparation only) and go through scipy like the reference.
def rotation_matrix_to_euler_angle(r: torch.Tensor, seq='XYZ'):
    from scipy.spatial.transform import Rotation
    rot = Rotation.from_matrix(r.clone().detach().cpu().view(-1, 3, 3).numpy())
    return torch.from_numpy(rot.as_euler(seq)).float().to(r.device)


def euler_angle_to_rotation_matrix(q: torch.Tensor, seq='XYZ'):
    from scipy.spatial.transform import Rotation
    rot = Rotation.from_euler(seq, q.clone().detach().cpu().view(-1, 3).numpy())
    return torch.from_numpy(rot.as_matrix()).float().to(q.device)


def rotation_matrix_to_euler_angle_np(r, seq='XYZ'):
    from scipy.spatial.transform import Rotation
    return Rotation.from_matrix(np.array(r).reshape(-1, 3, 3)).as_euler(seq)


def euler_angle_to_rotation_matrix_np(q, seq='XYZ'):
    from scipy.spatial.transform import Rotation
    return Rotation.from_euler(seq, np.array(q).reshape(-1, 3)).as_matrix()


def euler_convert_np(q, from_seq='XYZ', to_seq='XYZ'):
    from scipy.spatial.transform import Rotation
    return Rotation.from_euler(from_seq, q).as_euler(to_seq)


# ---------------------------------------------------------------------------------------------------------------
# spatial.py
def transformation_matrix_np(R, p):
    r"""spatial.py:19-35."""
    T = np.zeros((4, 4))
    T[:3, :3] = R
    T[:3, 3] = p
    T[3, 3] = 1
    return T


def adjoint_transformation_matrix_np(R, p):
    r"""spatial.py:38-57."""
    AdT = np.zeros((6, 6))
    AdT[:3, :3] = R
    AdT[3:, 3:] = R
    AdT[3:, :3] = np.dot(vector_cross_matrix_np(p), R)
    return AdT


def transformation_matrix(R: torch.Tensor, p: torch.Tensor):
    r"""[*,3,3], [*,3] -> [*,4,4]. spatial.py:60-75."""
    T = torch.zeros(list(R.shape[:-2]) + [4, 4], dtype=R.dtype, device=R.device)
    T[..., :3, :3] = R
    T[..., :3, 3] = p
    T[..., 3, 3] = 1
    return T


def decode_transformation_matrix(T: torch.Tensor):
    r"""spatial.py:78-87."""
    return T[..., :3, :3].clone(), T[..., :3, 3].clone()


def inverse_transformation_matrix(T: torch.Tensor):
    r"""spatial.py:90-101."""
    R, p = decode_transformation_matrix(T)
    invR = R.transpose(-1, -2)
    return transformation_matrix(invR, -torch.matmul(invR, p.unsqueeze(-1)).squeeze(-1))


def bone_vector_to_joint_position(bone_vec: torch.Tensor, parent):
    r"""[B, *] -> [B, J, 3]. spatial.py:126-145."""
    return _tree('rc_tree_bone_to_joint', bone_vec, parent, 3).view(bone_vec.shape[0], -1, 3)


def joint_position_to_bone_vector(joint_pos: torch.Tensor, parent):
    r"""[B, *] -> [B, J, 3]. spatial.py:148-167."""
    return _tree('rc_tree_joint_to_bone', joint_pos, parent, 3).view(joint_pos.shape[0], -1, 3)


def forward_kinematics_R(R_local: torch.Tensor, parent):
    r"""[B, *] -> [B, J, 3, 3]. spatial.py:170-194."""
    return _tree('rc_tree_fk_R', R_local, parent, 9).view(R_local.shape[0], -1, 3, 3)


def inverse_kinematics_R(R_global: torch.Tensor, parent):
    r"""[B, *] -> [B, J, 3, 3]. spatial.py:197-221."""
    return _tree('rc_tree_ik_R', R_global, parent, 9).view(R_global.shape[0], -1, 3, 3)


def forward_kinematics_T(T_local: torch.Tensor, parent):
    r"""[B, *] -> [B, J, 4, 4]. spatial.py:224-249."""
    return _tree('rc_tree_fk_T', T_local, parent, 16).view(T_local.shape[0], -1, 4, 4)


def inverse_kinematics_T(T_global: torch.Tensor, parent):
    r"""[B, *] -> [B, J, 4, 4]. spatial.py:252-277."""
    return _tree('rc_tree_ik_T', T_global, parent, 16).view(T_global.shape[0], -1, 4, 4)


def forward_kinematics(R_local: torch.Tensor, p_local: torch.Tensor, parent):
    r"""spatial.py:280-309."""
    R_local = R_local.view(R_local.shape[0], -1, 3, 3)
    p_local = p_local.view(p_local.shape[0], -1, 3)
    return decode_transformation_matrix(forward_kinematics_T(transformation_matrix(R_local, p_local), parent))
