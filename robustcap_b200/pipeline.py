r"""Callers / data formats on either side of the hot path, on the device.

* :func:`pack_rows` / :func:`load_aist_style` — what ``evaluate.py:24-52`` (dataset construction: camera-frame IMU data,
  pixel key points) and ``evaluate.py:68-73`` (per-sequence ``K^-1`` normalisation, camera-frame gravity) do with Python loops over
  ``RNNDataset`` lists, as ONE kernel that writes the packed ``[B, T_max, ...]`` tensors + ``lengths`` that
  :meth:`robustcap_b200.Net.forward_offline` consumes (``gravity=`` per row).  The on-disk schema is the reference's ``*.pt`` dict
  (``preprocess.py:50``: keys ``pose, tran, joint2d_mp, cam_K, cam_T, imu_ori, imu_acc, name``).
* :func:`synthesize_imu` — ``preprocess.py:22-33, 290-302``: synthetic IMU accelerations / orientations of a pose sequence
  (mesh FK of the six IMU vertices only, ``_syn_acc`` second differences).

Host code is tensor plumbing; the arithmetic is ``csrc/pipeline.cu`` behind ``rc_pack_inputs`` / ``rc_synthesize_imu``.
"""
from collections import namedtuple

import torch

from . import _lib
from . import math as M
from .constants import VI_MASK, JI_MASK

__all__ = ['PackedBatch', 'pack_rows', 'load_aist_style', 'synthesize_imu']

PackedBatch = namedtuple('PackedBatch', 'j2dc accc oric gravity lengths first_tran pose_t tran_t')


def _cat_offsets(seqs, dev, width):
    off = [0]
    for s in seqs:
        off.append(off[-1] + int(s.shape[0]))
    flat = torch.cat([s.reshape(s.shape[0], width).float() for s in seqs]).to(dev).contiguous()
    return flat, torch.tensor(off, dtype=torch.int64, device=dev)


@torch.no_grad()
def pack_rows(joint2d, imu_acc, imu_ori, cam_T, cam_K, src=None, image_size=(1920, 1080), tran=None, pose=None):
    r"""Pack dataset rows into network inputs (evaluate.py:38-52, 68-73).

    joint2d: list of B tensors ``[T_b, 33, 3]`` (u, v in [0, 1], confidence) — one per (sequence, camera) row.
    imu_acc / imu_ori: lists of S world-frame tensors ``[T_s, 6, 3]`` / ``[T_s, 6, 3, 3]``; ``src[b]`` = sequence of row b
    (default: row b uses sequence b).  cam_T ``[B,4,4]`` (Tcw), cam_K ``[B,3,3]``.
    tran / pose (optional, world-frame labels per sequence: ``[T_s,3]`` and axis-angle ``[T_s,24,3]``) give the camera-frame
    labels of evaluate.py:50-52 and ``first_tran`` (evaluate.py:74).
    Returns :class:`PackedBatch` on the current CUDA device."""
    lib = _lib.load()
    dev = _lib.require_cuda()
    B = len(joint2d)
    src = list(range(B)) if src is None else [int(s) for s in src]
    assert len(src) == B and len(imu_acc) == len(imu_ori)
    for b in range(B):
        assert joint2d[b].shape[0] == imu_acc[src[b]].shape[0] == imu_ori[src[b]].shape[0], 'row %d: lengths differ' % b
    Tmax = max(int(j.shape[0]) for j in joint2d)
    j2d, row_off = _cat_offsets(joint2d, dev, 99)
    acc, seq_off = _cat_offsets(imu_acc, dev, 18)
    ori, _ = _cat_offsets(imu_ori, dev, 54)
    cT = torch.as_tensor(cam_T).float().reshape(B, 16).to(dev).contiguous()
    cK = torch.as_tensor(cam_K).float().reshape(B, 9).to(dev).contiguous()
    dsrc = torch.tensor(src, dtype=torch.int32, device=dev)
    j2dc = torch.empty(B, Tmax, 33, 3, device=dev)
    accc = torch.empty(B, Tmax, 6, 3, device=dev)
    oric = torch.empty(B, Tmax, 6, 3, 3, device=dev)
    grav = torch.empty(B, 3, device=dev)
    lengths = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.check(lib.rc_pack_inputs(B, Tmax, _lib.dptr(dsrc), _lib.dptr(seq_off), _lib.dptr(row_off), _lib.dptr(j2d), _lib.dptr(acc),
                                  _lib.dptr(ori), _lib.dptr(cT), _lib.dptr(cK), float(image_size[0]), float(image_size[1]),
                                  _lib.dptr(j2dc), _lib.dptr(accc), _lib.dptr(oric), _lib.dptr(grav), _lib.dptr(lengths), _lib.stream()))
    first_tran = pose_t = tran_t = None
    if tran is not None:                                   # labels are not on the timed path: plain tensor ops (evaluate.py:50-52)
        Tc = cT.view(B, 4, 4)
        tran_t = [(Tc[b, :3, :3] @ tran[src[b]].to(dev).float().unsqueeze(-1)).squeeze(-1) + Tc[b, :3, 3] for b in range(B)]
        first_tran = torch.stack([t[0] for t in tran_t])
    if pose is not None:
        Tc = cT.view(B, 4, 4)
        pose_t = []
        for b in range(B):
            p = M.axis_angle_to_rotation_matrix(pose[src[b]].to(dev).float()).view(-1, 24, 3, 3)
            p[:, 0] = Tc[b, :3, :3] @ p[:, 0]
            pose_t.append(p)
    return PackedBatch(j2dc, accc, oric, grav, lengths, first_tran, pose_t, tran_t)


def load_aist_style(dataset, not_aligned=(), n_cam=9, image_size=(1920, 1080)):
    r"""The reference's AIST++ ``test.pt`` dict (preprocess.py:50) -> one packed batch with ``n_cam`` rows per sequence, the row
    order of evaluate.py:32-52.  Returns ``(PackedBatch, valid)`` with ``valid`` the row indices not in ``not_aligned``
    (evaluate.py:34-36)."""
    joint2d, cam_T, cam_K, src, valid = [], [], [], [], []
    for i in range(len(dataset['pose'])):
        for j in range(n_cam):
            if dataset['name'][i].replace('cAll', 'c0' + str(j + 1)) not in not_aligned:
                valid.append(len(src))
            joint2d.append(dataset['joint2d_mp'][i][j])
            cam_T.append(dataset['cam_T'][i][j])
            cam_K.append(dataset['cam_K'][i][j])
            src.append(i)
    batch = pack_rows(joint2d, dataset['imu_acc'], dataset['imu_ori'], torch.stack(cam_T), torch.stack(cam_K), src, image_size,
                      tran=dataset['tran'], pose=dataset['pose'])
    return batch, valid


@torch.no_grad()
def synthesize_imu(body_model, pose, tran=None, shape=None, smooth_n=2, vi_mask=VI_MASK, ji_mask=JI_MASK, return_aux=False):
    r"""Synthetic IMU readings of a pose sequence (preprocess.py:290-302): ``pose`` ``[T,24,3,3]`` local rotations, ``tran``
    ``[T,3]``, ``shape`` ``[10]`` or None (mean shape).  Returns ``imu_acc [T,6,3]`` (``_syn_acc`` of the vertices ``vi_mask``) and
    ``imu_ori [T,6,3,3]`` (global rotations of the joints ``ji_mask``); with ``return_aux`` also ``joint [T,24,3]`` and the IMU
    vertices ``[T,6,3]``."""
    lib = _lib.load()
    dev = _lib.require_cuda()
    p = pose.detach().reshape(-1, 24, 3, 3).to(dev, torch.float32).contiguous()
    n = p.shape[0]
    t = None if tran is None else tran.detach().reshape(n, 3).to(dev, torch.float32).contiguous()
    jrest = vrest = None
    if shape is not None:
        j, v = body_model.get_zero_pose_joint_and_vertex(shape.reshape(1, -1).to(body_model._v_template.device))
        jrest = j.reshape(1, 24, 3).expand(n, 24, 3).to(dev, torch.float32).contiguous()
        vrest = v.reshape(-1, 3)[list(vi_mask)].to(dev, torch.float32).contiguous()
    vi = torch.tensor(list(vi_mask), dtype=torch.int32)
    ji = torch.tensor(list(ji_mask), dtype=torch.int32)
    k = len(vi)
    acc = torch.empty(n, k, 3, device=dev)
    ori = torch.empty(n, k, 3, 3, device=dev)
    joint = torch.empty(n, 24, 3, device=dev)
    vimu = torch.empty(n, k, 3, device=dev)
    _lib.check(lib.rc_synthesize_imu(body_model._native(), _lib.dptr(p), _lib.dptr(t), _lib.dptr(jrest), _lib.dptr(vrest),
                                     _lib.hptr(vi), _lib.hptr(ji), k, int(smooth_n), n, _lib.dptr(acc), _lib.dptr(ori),
                                     _lib.dptr(joint), _lib.dptr(vimu), _lib.stream()))
    if return_aux:
        return acc, ori, joint, vimu
    return acc, ori
