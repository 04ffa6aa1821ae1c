"""SURVEY 8(f) row 4 — live wire formats.  CPU: the native parser / formatter (host code of the C-ABI library) against strings built
by the reference's own statements (tests/golden/live.json).  GPU: LiveSession.feed (datagram -> forward_online -> Unity message)
against the CPU oracle post-processed like live_server.py:49-58."""
import json
import os

import pytest
import torch

from robustcap_b200 import synthetic


@pytest.fixture(scope='module')
def live(golden_dir):
    from robustcap_b200 import _lib
    _lib.build()
    return json.load(open(os.path.join(golden_dir, 'live.json')))


def test_parse_frame_matches_reference(live):
    from robustcap_b200.live import parse_frame
    for f in live:
        uv, ori, acc, rcm = parse_frame(f['datagram'].encode())
        assert torch.equal(uv.reshape(-1), torch.tensor(f['uv'])) and torch.equal(ori.reshape(-1), torch.tensor(f['ori']))
        assert torch.equal(acc.reshape(-1), torch.tensor(f['acc'])) and torch.equal(rcm.reshape(-1), torch.tensor(f['rcm']))


def test_parse_frame_rejects_malformed(live):
    from robustcap_b200.live import parse_frame
    good = live[0]['datagram']
    for bad in (good.replace('#', ',', 1), good[: good.index('#')], good.replace(',', ',x', 1), good + ',1.0'):
        with pytest.raises(RuntimeError):
            parse_frame(bad.encode())


def test_format_pose_matches_reference(live):
    from robustcap_b200.live import format_pose
    for f in live:
        msg = format_pose(torch.tensor(f['pose_aa']), torch.tensor(f['tran']))
        assert msg.decode() == f['unity']


def test_parse_imu_packet_matches_reference(live):
    """live_demo_sync.py:262-268 (get_from_udp), statement for statement."""
    import numpy as np
    from robustcap_b200.live import parse_imu_packet
    N = 6
    raw = np.random.RandomState(3).randn(8 * N).astype(np.float32).tobytes()
    data = np.frombuffer(raw, np.float32).copy()
    t, q, a = data[:N], data[N:5 * N].reshape(N, 4), data[5 * N:].reshape(N, 3)
    gt, gq, ga = parse_imu_packet(raw, N)
    assert gt == t.tolist() and torch.equal(gq, torch.from_numpy(q)) and torch.equal(ga, torch.from_numpy(a))
    with pytest.raises(RuntimeError):
        parse_imu_packet(raw[:-4], N)


@pytest.mark.gpu
def test_live_session_vs_oracle(live, assets):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import robustcap_b200 as rb
    from robustcap_b200.live import LiveSession
    from oracle.kinematics import BodyOracle
    from oracle.fusion import FusionOracle
    from oracle.rotations import matrix_to_axis_angle
    sd = synthetic.make_state_dict(0, 'contact')
    net = rb.Net(rb.ParametricModel(assets['smpl_file']))
    net.load_state_dict(sd)
    inp = synthetic.make_inputs(1, 10, seed=9, conf='mixed')
    RCM = synthetic._random_rotations(1, torch.Generator().manual_seed(1))[0]
    fmt = lambda x: ','.join(str(v) for v in x.numpy().reshape(-1))
    sess = LiveSession(net)
    msgs = []
    mk = lambda t: '#'.join((fmt(inp['j2dc'][0, t]), fmt(inp['oric'][0, t]), fmt(inp['accc'][0, t]), fmt(RCM))).encode()
    class_gravity = rb.Net.gravityc.clone()
    assert sess.feed(mk(0)) is None          # the first datagram only calibrates (live_server.py:32-35), like the reference
    assert torch.equal(rb.Net.gravityc, class_gravity) and 'gravityc' in net.__dict__       # gravity lands on the instance, not the class
    for t in range(10):
        msgs.append(sess.feed(mk(t)).decode())
    grav = RCM @ torch.tensor([0., -1, 0.])
    o = FusionOracle(sd, BodyOracle(assets['smpl_file']))
    op, ot = o.run(inp['j2dc'][0], inp['accc'][0], inp['oric'][0], first_frame=True, gravity=grav)
    stran = None
    for t in range(10):
        pose = op[t].clone()
        pose[0] = RCM.T @ pose[0]
        tran = RCM.T @ ot[t]
        stran = tran.clone() if stran is None else stran
        aa = matrix_to_axis_angle(pose).reshape(-1)
        ps, ts = msgs[t].rstrip('$').split('#')
        got_p = torch.tensor([float(v) for v in ps.split(',')])
        got_t = torch.tensor([float(v) for v in ts.split(',')])
        assert got_p.numel() == 72 and got_t.numel() == 3 and msgs[t].endswith('$')
        assert (got_p - aa).abs().max().item() < 2e-4, t           # 1e-4 rad of the float path + 6 significant digits of %g
        assert (got_t - (tran - stran)).abs().max().item() < 1e-3, t
