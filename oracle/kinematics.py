"""Oracle (test infrastructure): kinematic-tree math and the SMPL body model.  See ``oracle/__init__.py``.

Restates ``articulate/math/spatial.py`` and the compute part of ``articulate/model.py``.
"""
import pickle

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------------------------
# SE(3) helpers (spatial.py:60-101)
def make_T(R, p):
    """spatial.py:60-75: [[R, p], [0, 1]]."""
    top = torch.cat((R, p.unsqueeze(-1)), dim=-1)
    bottom = torch.zeros(list(top.shape[:-2]) + [1, 4], dtype=R.dtype)
    bottom[..., 0, 3] = 1
    return torch.cat((top, bottom), dim=-2)


def split_T(T):
    """spatial.py:78-87."""
    return T[..., :3, :3].clone(), T[..., :3, 3].clone()


def invert_T(T):
    """spatial.py:90-101: [[R^T, -R^T p], [0, 1]]."""
    R, p = split_T(T)
    Rt = R.transpose(-1, -2)
    return make_T(Rt, -torch.matmul(Rt, p.unsqueeze(-1)).squeeze(-1))


# ----------------------------------------------------------------------------------------------------------------
# tree prefix products / sums (spatial.py:104-123); parent[i] < i, parent[0] is the base
def tree_accumulate(local, parent, op):
    """spatial.py:104-112: global[i] = op(global[parent[i]], local[i])."""
    out = [local[:, 0]]
    for i in range(1, len(parent)):
        out.append(op(out[parent[i]], local[:, i]))
    return torch.stack(out, dim=1)


def tree_differentiate(glob, parent, op, inv):
    """spatial.py:115-123: local[i] = op(inv(global[parent[i]]), global[i])."""
    out = [glob[:, 0]]
    for i in range(1, len(parent)):
        out.append(op(inv(glob[:, parent[i]]), glob[:, i]))
    return torch.stack(out, dim=1)


def bone_to_joint(bone, parent):
    """spatial.py:126-145."""
    return tree_accumulate(bone.reshape(bone.shape[0], -1, 3), parent, torch.add)


def joint_to_bone(joint, parent):
    """spatial.py:148-167: bone[i] = -joint[parent] + joint[i]."""
    return tree_differentiate(joint.reshape(joint.shape[0], -1, 3), parent, torch.add, torch.neg)


def fk_R(R_local, parent):
    """spatial.py:170-194."""
    return tree_accumulate(R_local.reshape(R_local.shape[0], -1, 3, 3), parent, torch.bmm)


def ik_R(R_global, parent):
    """spatial.py:197-221: local[i] = global[parent]^T @ global[i]."""
    return tree_differentiate(R_global.reshape(R_global.shape[0], -1, 3, 3), parent, torch.bmm,
                              lambda m: m.transpose(1, 2))


def fk_T(T_local, parent):
    """spatial.py:224-249."""
    return tree_accumulate(T_local.reshape(T_local.shape[0], -1, 4, 4), parent, torch.bmm)


def ik_T(T_global, parent):
    """spatial.py:252-277."""
    return tree_differentiate(T_global.reshape(T_global.shape[0], -1, 4, 4), parent, torch.bmm, invert_T)


def bbox_scale(uv):
    """net/sig_mp.py:277-284: max(range of u, range of v) over the keypoint axis."""
    u, v = uv[..., 0], uv[..., 1]
    return torch.max(u.max(dim=-1).values - u.min(dim=-1).values, v.max(dim=-1).values - v.min(dim=-1).values)


# ----------------------------------------------------------------------------------------------------------------
class BodyOracle:
    """articulate/model.py:17-241 (compute part): SMPL constants + FK / LBS."""

    def __init__(self, model_file, dtype=torch.float32):
        with open(model_file, 'rb') as f:
            d = pickle.load(f, encoding='latin1')
        t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64)).float().to(dtype)  # model.py:31-36 (.float())
        self.dtype = dtype
        self.J_regressor = t(d['J_regressor'].toarray())
        self.skin_w = t(d['weights'])
        self.posedirs = t(d['posedirs'])
        self.shapedirs = t(np.array(d['shapedirs']))
        self.v_template = t(d['v_template'])
        self.J = t(d['J'])
        self.parent = d['kintree_table'][0].tolist()
        self.parent[0] = None

    def zero_pose(self, shape=None):
        """model.py:78-93."""
        if shape is None:
            return self.J - self.J[:1], self.v_template - self.J[:1]
        shape = shape.reshape(-1, 10)
        v = torch.tensordot(shape, self.shapedirs, dims=([1], [2])) + self.v_template
        j = torch.matmul(self.J_regressor, v)
        return j - j[:, :1], v - j[:, :1]

    def bones(self, shape=None):
        j, _ = self.zero_pose(shape)
        if j.dim() == 2:
            j = j.unsqueeze(0)
        return joint_to_bone(j, self.parent)

    def forward_kinematics(self, pose, shape=None, tran=None, calc_mesh=False, use_pose_blendshape=False):
        """model.py:209-241."""
        pose = pose.reshape(pose.shape[0], -1, 3, 3)
        B = pose.shape[0]
        j, v = [x.expand(B, -1, -1) for x in self.zero_pose(shape)]
        Tl = make_T(pose, joint_to_bone(j, self.parent))
        Tg = fk_T(Tl, self.parent)
        Rg, pg = split_T(Tg)
        add = (lambda x: x) if tran is None else (lambda x: x + tran.reshape(-1, 1, 3))
        if not calc_mesh:
            return Rg, add(pg)
        j0 = torch.cat((j, torch.zeros_like(j[..., :1])), dim=-1).unsqueeze(-1)
        Tg = Tg.clone()
        Tg[..., -1:] = Tg[..., -1:] - torch.matmul(Tg, j0)
        Tv = torch.tensordot(Tg, self.skin_w, dims=([1], [1])).permute(0, 3, 1, 2)
        if use_pose_blendshape:
            r = (pose[:, 1:] - torch.eye(3, dtype=pose.dtype)).flatten(1)
            v = v + torch.tensordot(r, self.posedirs, dims=([1], [2]))
        v1 = torch.cat((v, torch.ones_like(v[..., :1])), dim=-1).unsqueeze(-1)
        vg = torch.matmul(Tv, v1).squeeze(-1)[..., :3]
        return Rg, add(pg), add(vg)


def mediapipe_points(vert, joint, mp_mask):
    """net/sig_mp.py:287-299 / utils.py:129-135 (batched over the leading axis)."""
    out = vert[:, mp_mask].clone()
    out[:, 11:17] = joint[:, 16:22]
    out[:, 23:25] = joint[:, 1:3]
    out[:, 25:27] = joint[:, 4:6]
    out[:, 27:29] = joint[:, 7:9]
    return out
