r"""Live wire formats: the per-frame loop body of ``live_server.py:40-59`` without the sockets.

``LiveSession.feed(datagram)`` takes the UDP text frame of ``live_detector.py:57-61`` (``uv#ori#acc#RCM``), runs
``Net.forward_online`` (the first frame with ``first_frame=True``), rotates the result back with ``RCM^T``, subtracts the first
translation, converts to axis-angle and returns the TCP message Unity expects (``pose#tran$``).  Parsing and formatting are native
host code (``rc_live_parse_frame`` / ``rc_live_format_pose``); the caller owns the sockets.
"""
import ctypes

import torch

from . import _lib
from . import math as M

__all__ = ['parse_frame', 'format_pose', 'parse_imu_packet', 'LiveSession']


def parse_frame(datagram: bytes):
    """``uv [33,3], ori [6,3,3], acc [6,3], RCM [3,3]`` float32 CPU tensors (live_server.py:42-45)."""
    lib = _lib.load()
    uv, ori, acc, rcm = torch.empty(33, 3), torch.empty(6, 3, 3), torch.empty(6, 3), torch.empty(3, 3)
    _lib.check(lib.rc_live_parse_frame(datagram, len(datagram), _lib.hptr(uv), _lib.hptr(ori), _lib.hptr(acc), _lib.hptr(rcm)))
    return uv, ori, acc, rcm


def format_pose(pose_aa: torch.Tensor, tran: torch.Tensor) -> bytes:
    """live_server.py:57-58."""
    lib = _lib.load()
    p = pose_aa.detach().reshape(72).to('cpu', torch.float32).contiguous()
    t = tran.detach().reshape(3).to('cpu', torch.float32).contiguous()
    buf = ctypes.create_string_buffer(2048)
    n = lib.rc_live_format_pose(_lib.hptr(p), _lib.hptr(t), buf, 2048)
    if n < 0:
        _lib.check(n)
    return buf.raw[:n]


def parse_imu_packet(data: bytes, n_imu: int):
    """``t`` (list of n floats), ``q [n,4]`` (wxyz), ``a [n,3]`` of the sensor server's binary datagram (live_demo_sync.py:262-268)."""
    lib = _lib.load()
    t, q, a = torch.empty(n_imu), torch.empty(n_imu, 4), torch.empty(n_imu, 3)
    _lib.check(lib.rc_live_parse_imu_packet(data, len(data), int(n_imu), _lib.hptr(t), _lib.hptr(q), _lib.hptr(a)))
    return t.tolist(), q, a


class LiveSession:
    def __init__(self, net):
        self.net = net
        self.RCM = None
        self.stran = None

    def feed(self, datagram: bytes):
        """Returns the Unity message, or ``None`` for the very first datagram: ``run_live_demo`` only reads the calibration
        ``RCM`` (and sets the instance's gravity) from it and starts inference with the next one (live_server.py:32-35)."""
        uv, ori, acc, rcm = parse_frame(datagram)
        if self.RCM is None:                                                    # live_server.py:32-35
            self.RCM = rcm
            self.net.gravityc = torch.matmul(rcm, torch.tensor([0., -1, 0.]).unsqueeze(-1)).squeeze(-1)   # on the instance
            return None
        pose, tran = self.net.forward_online(uv, acc, ori, first_frame=self.stran is None)   # :46-48
        pose = pose.clone()
        pose[0] = self.RCM.T.matmul(pose[0])                                    # :49-50
        tran = self.RCM.T.matmul(tran.unsqueeze(-1)).squeeze(-1)                # :51
        if self.stran is None:
            self.stran = tran.clone()
        tran = tran - self.stran                                                # :52-54
        aa = M.rotation_matrix_to_axis_angle(pose).view(-1)                     # :55
        return format_pose(aa, tran)
