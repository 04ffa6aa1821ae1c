r"""``articulate.ParametricModel`` (reference ``articulate/model.py:17-241``) on the B200 library.

Loads the official SMPL pickle exactly like the reference, keeps the constants as torch tensors on the requested
device for API compatibility (``_J``, ``_v_template``, ``parent`` ...) and owns an ``rc_model`` handle with the same
constants resident in HBM for the kernels.  Visualisation / export helpers of the reference are out of scope.
"""
import pickle

import numpy as np
import torch

from . import _lib
from . import math as M
from .constants import MP_MASK

__all__ = ['ParametricModel']


class ParametricModel:
    def __init__(self, official_model_file: str, use_pose_blendshape=False, device=torch.device('cpu')):
        r"""model.py:21-40."""
        with open(official_model_file, 'rb') as f:
            data = pickle.load(f, encoding='latin1')
        self._J_regressor = torch.from_numpy(data['J_regressor'].toarray()).float().to(device)
        self._skinning_weights = torch.from_numpy(np.asarray(data['weights'])).float().to(device)
        self._posedirs = torch.from_numpy(np.asarray(data['posedirs'])).float().to(device)
        self._shapedirs = torch.from_numpy(np.array(data['shapedirs'])).float().to(device)
        self._v_template = torch.from_numpy(np.asarray(data['v_template'])).float().to(device)
        self._J = torch.from_numpy(np.asarray(data['J'])).float().to(device)
        self.face = data['f']
        self.parent = data['kintree_table'][0].tolist()
        self.parent[0] = None
        self.use_pose_blendshape = use_pose_blendshape
        self._handle = None
        self._dev_consts = None

    # ---- native handle ---------------------------------------------------------------------------------------
    def _native(self):
        if self._handle is None:
            lib = _lib.load()
            _lib.require_cuda()
            j, v = getattr(self, '_shaped_jv', None) or self.get_zero_pose_joint_and_vertex()
            j = j.detach().cpu().float().contiguous()
            v = v.detach().cpu().float().contiguous()
            w = self._skinning_weights.detach().cpu().float().contiguous()
            par = torch.tensor([-1] + [int(p) for p in self.parent[1:]], dtype=torch.int32)
            mp = torch.tensor(MP_MASK, dtype=torch.int32)
            h = _lib.vp()
            _lib.check(lib.rc_model_create(_lib.ctypes.byref(h), _lib.hptr(j), _lib.hptr(v), _lib.hptr(w), v.shape[0],
                                           _lib.hptr(par), _lib.hptr(mp)))
            self._handle = h
        return self._handle

    def shaped(self, shape: torch.Tensor):
        r"""A view of this body with the rest pose of ``shape`` ([10] betas) baked into the native constants — what
        ``forward_kinematics(..., shape=shape)`` uses in the reference (model.py:78-93, 224-241) — for the native paths that take a
        body handle (SMPLify with ``shape=``, temporal_smplify.py:84-86, 158-159).  Shares all tensors with ``self``."""
        import copy
        other = copy.copy(self)
        other._handle = None
        other._dev_consts = None
        j, v = self.get_zero_pose_joint_and_vertex(shape.detach().reshape(1, 10).to(self._v_template.device, torch.float32))
        other._shaped_jv = (j.reshape(24, 3), v.reshape(-1, 3))
        return other

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().rc_model_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # ---- reference API -----------------------------------------------------------------------------------------
    def get_zero_pose_joint_and_vertex(self, shape: torch.Tensor = None):
        r"""model.py:78-93 (blend-shape GEMMs are plain library matmuls)."""
        if shape is None:
            return self._J - self._J[:1], self._v_template - self._J[:1]
        shape = shape.view(-1, 10)
        v = torch.tensordot(shape, self._shapedirs, dims=([1], [2])) + self._v_template
        j = torch.matmul(self._J_regressor, v)
        return j - j[:, :1], v - j[:, :1]

    def bone_vector_to_joint_position(self, bone_vec: torch.Tensor):
        return M.bone_vector_to_joint_position(bone_vec, self.parent)

    def joint_position_to_bone_vector(self, joint_pos: torch.Tensor):
        return M.joint_position_to_bone_vector(joint_pos, self.parent)

    def forward_kinematics_R(self, R_local: torch.Tensor):
        return M.forward_kinematics_R(R_local, self.parent)

    def inverse_kinematics_R(self, R_global: torch.Tensor):
        return M.inverse_kinematics_R(R_global, self.parent)

    def forward_kinematics_T(self, T_local: torch.Tensor):
        return M.forward_kinematics_T(T_local, self.parent)

    def inverse_kinematics_T(self, T_global: torch.Tensor):
        return M.inverse_kinematics_T(T_global, self.parent)

    def forward_kinematics(self, pose: torch.Tensor, shape: torch.Tensor = None, tran: torch.Tensor = None,
                           calc_mesh=False):
        r"""model.py:209-241: global joint rotations, joint positions and (optionally) the skinned mesh."""
        lib = _lib.load()
        dev = _lib.require_cuda()
        out_dev = pose.device
        B = pose.shape[0]
        p = pose.detach().reshape(B, -1).to(device=dev, dtype=torch.float32).contiguous()
        assert p.shape[1] == 216, 'SMPL pose must reshape to [B, 24, 3, 3]'
        t = None if tran is None else tran.detach().reshape(-1, 3).to(device=dev, dtype=torch.float32).expand(B, 3).contiguous()
        jr = vr = None
        if shape is not None:
            j, v = self.get_zero_pose_joint_and_vertex(shape.to(self._shapedirs.device))
            jr = j.to(device=dev, dtype=torch.float32).expand(B, -1, -1).contiguous()
            vr = v.to(device=dev, dtype=torch.float32).expand(B, -1, -1).contiguous()
        if self.use_pose_blendshape and calc_mesh:
            r = (p.view(B, 24, 3, 3)[:, 1:] - torch.eye(3, device=dev)).flatten(1)
            base = vr if vr is not None else self.get_zero_pose_joint_and_vertex()[1].to(dev).expand(B, -1, -1)
            vr = (base + torch.tensordot(r, self._posedirs.to(dev), dims=([1], [2]))).contiguous()
        nv = self._v_template.shape[0]
        Rg = torch.empty(B, 24, 3, 3, device=dev)
        joint = torch.empty(B, 24, 3, device=dev)
        vert = torch.empty(B, nv, 3, device=dev) if calc_mesh else None
        _lib.check(lib.rc_model_forward_kinematics(self._native(), _lib.dptr(p), _lib.dptr(t), _lib.dptr(jr), _lib.dptr(vr),
                                                   B, _lib.dptr(Rg), _lib.dptr(joint), _lib.dptr(vert), _lib.stream()))
        if not calc_mesh:
            return Rg.to(out_dev), joint.to(out_dev)
        return Rg.to(out_dev), joint.to(out_dev), vert.to(out_dev)

    def keypoints33(self, pose: torch.Tensor, tran: torch.Tensor = None):
        r"""The 33 synthetic MediaPipe points of ``sync_mp3d`` (net/sig_mp.py:287-299) without skinning the full mesh:
        returns (joint [B,24,3], keypoints [B,33,3])."""
        lib = _lib.load()
        dev = _lib.require_cuda()
        B = pose.shape[0]
        p = pose.detach().reshape(B, 216).to(device=dev, dtype=torch.float32).contiguous()
        t = None if tran is None else tran.detach().reshape(-1, 3).to(device=dev, dtype=torch.float32).expand(B, 3).contiguous()
        joint = torch.empty(B, 24, 3, device=dev)
        kp = torch.empty(B, 33, 3, device=dev)
        _lib.check(lib.rc_model_keypoints(self._native(), _lib.dptr(p), _lib.dptr(t), B, _lib.dptr(joint), _lib.dptr(kp), _lib.stream()))
        return joint.to(pose.device), kp.to(pose.device)

    # visualisation / export of the reference (model.py:42-76, 243-418) is out of scope
    def _unsupported(self, *a, **k):
        raise NotImplementedError('visualisation/export helpers of articulate.ParametricModel are out of scope of the B200 hot path')

    save_obj_mesh = save_unity_motion = view_joint = view_mesh = view_motion = view_mesh_overlay = _unsupported
