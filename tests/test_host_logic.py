"""CPU check of the scalar device logic (rc_rows.h / rc_math.h / rc_pack.h) through tests/host_harness.cpp.

The harness is TEST CODE: it compiles the same host/device functions the CUDA kernels call, plus a naive CPU
restatement of the packed-layout LSTM, and replays fusion.cu's per-frame pass order.  It exists because the build
container has no GPU; the product library has no CPU path.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import kinematics as K
from oracle import rotations as R
from robustcap_b200 import synthetic
from robustcap_b200.constants import MP_MASK, NET_ORDER
from test_oracle_golden import ONLINE_CASES, load, pose_angle, get_sd

HERE = os.path.dirname(os.path.abspath(__file__))
FP = ctypes.POINTER(ctypes.c_float)


def fp(t):
    assert t.dtype == torch.float32 and t.is_contiguous()
    return ctypes.cast(t.data_ptr(), FP)


@pytest.fixture(scope='module')
def lib():
    out = os.path.join(HERE, '_build')
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, 'libhost_harness.so')
    src = os.path.join(HERE, 'host_harness.cpp')
    deps = [src] + [os.path.join(HERE, '..', 'robustcap_b200', 'csrc', f) for f in ('rc_rows.h', 'rc_math.h', 'rc_pack.h')]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', so, src])
    L = ctypes.CDLL(so)
    L.hh_create.restype = ctypes.c_void_p
    L.hh_conf_mean.restype = ctypes.c_float
    return L


def make_harness(L, body, sd, live=0):
    j0, v0 = body.zero_pose()
    parent = torch.tensor([-1] + body.parent[1:], dtype=torch.int32)
    mp = torch.tensor(MP_MASK, dtype=torch.int32)
    h = ctypes.c_void_p(L.hh_create(fp(j0.contiguous()), fp(v0.contiguous()), fp(body.skin_w.contiguous()), v0.shape[0],
                                    ctypes.cast(parent.data_ptr(), ctypes.POINTER(ctypes.c_int)),
                                    ctypes.cast(mp.data_ptr(), ctypes.POINTER(ctypes.c_int)), live))
    for ni, name in enumerate(NET_ORDER):
        keys = ['linear1.weight', 'linear1.bias']
        for l in range(2):
            keys += ['rnn.weight_ih_l%d' % l, 'rnn.weight_hh_l%d' % l, 'rnn.bias_ih_l%d' % l, 'rnn.bias_hh_l%d' % l]
        keys += ['linear2.weight', 'linear2.bias']
        L.hh_set_net(h, ni, *[fp(sd[name + '.' + k].contiguous()) for k in keys])
    L.hh_set_init(h, *[fp(sd['rnn2.init_net.%d.%s' % (k, w)].contiguous()) for k in (0, 2, 4) for w in ('weight', 'bias')])
    return h


def test_rotation_math(lib, golden_dir):
    g = load(golden_dir, 'math.npz')

    def run(fn, x, width):
        x = x.contiguous()
        o = torch.empty(x.shape[0], width)
        fn(fp(x), fp(o), x.shape[0])
        return o

    assert (run(lib.hh_r6d_to_mat, g['r6d'], 9).view(-1, 3, 3) - g['r6d_to_R']).abs().max() < 1e-6
    assert (run(lib.hh_aa_to_mat, g['aa'], 9).view(-1, 3, 3) - g['aa_to_R']).abs().max() < 1e-6
    assert (run(lib.hh_batch_rodrigues, g['aa'], 9).view(-1, 3, 3) - g['batch_rodrigues']).abs().max() < 1e-6
    assert (run(lib.hh_mat_to_aa, g['R_noisy'].reshape(-1, 9), 3) - g['R_to_aa']).abs().max() < 2e-6
    assert (run(lib.hh_quat_to_mat, g['q'], 9).view(-1, 3, 3) - g['q_to_R']).abs().max() < 1e-6
    assert (run(lib.hh_quat_to_aa, g['q'], 3) - g['q_to_aa']).abs().max() < 1e-5
    assert (run(lib.hh_aa_to_quat, g['aa'], 4) - g['aa_to_q']).abs().max() < 1e-6
    o = torch.empty(64, 4)
    lib.hh_quat_mul(fp(g['q'].contiguous()), fp(g['q2'].contiguous()), fp(o), 64)
    assert (o - g['q_prod']).abs().max() < 1e-6
    # far-from-orthonormal inputs: the Newton polar projection must agree with cv2's SVD projection
    gen = torch.Generator().manual_seed(5)
    bad = synthetic._random_rotations(256, gen) * (0.5 + torch.rand(256, 1, 1, generator=gen)) + 0.2 * torch.randn(256, 3, 3, generator=gen)
    ref = R.matrix_to_axis_angle_cv2(bad)
    assert (run(lib.hh_mat_to_aa, bad.reshape(-1, 9), 3) - ref).abs().max() < 5e-6


def test_conf_mean_matches_torch(lib):
    gen = torch.Generator().manual_seed(3)
    for _ in range(500):
        kp = torch.rand(33, 3, generator=gen)
        assert lib.hh_conf_mean(fp(kp)) == kp[:, -1].mean().item()


def test_fk_keypoints(lib, golden_dir, assets):
    g = load(golden_dir, 'kinematics.npz')
    body = K.BodyOracle(assets['smpl_file'])
    h = make_harness(lib, body, get_sd(0, 'default'))
    for b in range(g['pose'].shape[0]):
        joint, kp = torch.empty(24, 3), torch.empty(33, 3)
        lib.hh_fk_keypoints(h, fp(g['pose'][b].contiguous()), fp(g['tran'][b].contiguous()), fp(joint), fp(kp))
        assert (joint - g['fk_mesh_joint'][b]).abs().max() < 2e-6
        assert (kp - g['fk_mesh_vert_mp'][b]).abs().max() < 2e-6
    lib.hh_destroy(h)


@pytest.mark.parametrize('case', ONLINE_CASES, ids=[c[0] for c in ONLINE_CASES])
def test_online_logic(lib, golden_dir, assets, case):
    name, wseed, variant, conf, iseed, start, Tn = case
    Tn = min(Tn, 40) if not name.startswith('live_') else Tn
    g = load(golden_dir, 'online_%s.npz' % name)
    body = K.BodyOracle(assets['smpl_file'])
    h = make_harness(lib, body, get_sd(wseed, variant), live=int(name.startswith('live_')))
    lib.hh_set_gravity(h, fp(g['gravity'].contiguous()))
    lib.hh_reset(h)
    poses, trans = torch.empty(Tn, 24, 3, 3), torch.empty(Tn, 3)
    ft = torch.tensor([0., 0., 4.])
    for t in range(Tn):
        flags = 0
        if t == 0 and start == 'first_frame':
            flags = 1
        if t == 0 and start == 'first_tran':
            flags = 2
        lib.hh_step(h, fp(g['j2dc'][t].contiguous()), fp(g['accc'][t].contiguous()), fp(g['oric'][t].contiguous()), flags,
                    fp(ft), fp(poses[t]), fp(trans[t]))
    ang = pose_angle(poses, g['pose'][:Tn]).max().item()
    terr = (trans - g['tran'][:Tn]).abs().max().item()
    lib.hh_destroy(h)
    assert ang < 6e-5, ang
    assert terr < 2e-5, terr
