"""robustcap_b200 — B200-native implementation of RobustCap's per-frame fusion + kinematics hot path.

Public surface mirrors the reference (shaohua-pan/RobustCap):
  robustcap_b200.math            == articulate.math
  robustcap_b200.ParametricModel == articulate.ParametricModel
  robustcap_b200.RNN / RNNWithInit
  robustcap_b200.Net             == net.sig_mp.Net (+ forward_offline)
  robustcap_b200.smplify_runner  == net.smplify.run.smplify_runner
  robustcap_b200.pipeline        dataset rows -> packed device batch (evaluate.py:24-73), IMU synthesis (preprocess.py:22-33)
  robustcap_b200.metrics         cal_mpjpe on the device (evaluate.py:120-133)
Drop-in module names (``articulate``, ``net.sig_mp``, ``net.smplify.run``, ``config``, ``utils``) live in
``robustcap_b200/dropin`` — put that directory first on ``sys.path``.
"""
from . import constants
from . import math
from .model import ParametricModel
from .rnn import RNN, RNNWithInit
from .net import Net, get_bbox_scale, sync_mp3d

__all__ = ['constants', 'math', 'ParametricModel', 'RNN', 'RNNWithInit', 'Net', 'get_bbox_scale', 'sync_mp3d',
           'smplify_runner', 'TemporalSMPLify']


def __getattr__(name):
    if name in ('smplify_runner', 'TemporalSMPLify'):
        from . import smplify
        return getattr(smplify, name)
    raise AttributeError(name)
