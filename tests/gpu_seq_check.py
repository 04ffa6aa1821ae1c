"""GPU aid (not a test): persistent sequence kernel (gemm mode 3) vs the grouped kernel (mode 2): results and timing.
Usage: python tests/gpu_seq_check.py [B] [T] [conf] [reps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import robustcap_b200 as rb
from robustcap_b200 import synthetic
from test_oracle_golden import pose_angle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
conf = sys.argv[3] if len(sys.argv) > 3 else 'mixed'
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
net = rb.Net(rb.ParametricModel(assets['smpl_file']))
net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
inp = synthetic.make_inputs(B, T, seed=1000, conf=conf)
rb.Net.gravityc = inp['gravity'].clone()
j, a, o = inp['j2dc'].cuda(), inp['accc'].cuda(), inp['oric'].cuda()
ft = torch.tensor([0., 0., 4.], device='cuda')
res = {}
for mode in (2, 3):
    net.set_gemm_mode(mode)
    log = torch.zeros(B, T, dtype=torch.int32, device='cuda')
    p, t = net.forward_offline(j, a, o, first_tran=ft, branch_log=log)
    torch.cuda.synchronize()
    res[mode] = (p.cpu(), t.cpu(), log.cpu())
    for _ in range(2):
        net.forward_offline(j, a, o, first_tran=ft)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        net.forward_offline(j, a, o, first_tran=ft)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('mode %d: %.2f ms per pass, %.1f us per frame, %.0f frames/s' % (mode, ms, ms / T * 1e3, B * T / ms * 1e3), flush=True)
p2, t2, l2 = res[2]
p3, t3, l3 = res[3]
ang = pose_angle(p2, p3).view(B, T, 24).amax(dim=2)
dt = (t2 - t3).abs().amax(dim=2)
print('mode 3 vs mode 2: max pose angle %.3e rad, max tran %.3e m, branch logs equal: %s, finite: %s'
      % (ang.max().item(), dt.max().item(), bool(torch.equal(l2, l3)), bool(torch.isfinite(p3).all() and torch.isfinite(t3).all())))
per_t = ang.amax(dim=0)
print('per-frame max angle (first 12):', ' '.join('%.1e' % v for v in per_t[:12].tolist()))
bad = (ang > 1e-4).nonzero()
if len(bad):
    print('first mismatching (row, frame):', bad[:8].tolist())
    print('zero pose rows in mode 3:', int((p3.abs().sum(dim=(2, 3, 4)) == 0).sum()))
if os.environ.get('RC_SEQ_STATS'):
    import ctypes
    from robustcap_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_longlong * (36 * 4))()
    n = lib.rc_state_debug_seq_stats(net._states[B], buf, 36)
    for jn in range(n):
        print('job %2d: tiles %8d  dep1 wait %12d clk  dep2 wait %12d clk  skipped %8d' % (jn, buf[jn * 4], buf[jn * 4 + 1], buf[jn * 4 + 2], buf[jn * 4 + 3]))
