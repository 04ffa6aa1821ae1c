"""GPU profiling aid (not a test): tile traces of the persistent grouped tcgen05 kernel (gemm mode 2).
Usage: python tests/gpu_phase_trace.py [B] [T] [conf]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robustcap_b200 as rb
from robustcap_b200 import synthetic, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
conf = sys.argv[3] if len(sys.argv) > 3 else 'mixed'
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
net = rb.Net(rb.ParametricModel(assets['smpl_file']))
net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
inp = synthetic.make_inputs(B, T, seed=31, conf=conf)
rb.Net.gravityc = inp['gravity'].clone()
net.set_gemm_mode(2)
j, a, o = inp['j2dc'].cuda(), inp['accc'].cuda(), inp['oric'].cuda()
ft = torch.tensor([0., 0., 4.])
net.forward_offline(j, a, o, first_tran=ft, use_graph=False)
lib = _lib.load()
st = net._state(B)
_lib.check(lib.rc_state_debug_phase_trace(st, 1, 0, None, 0, None))
net.forward_offline(j[:, -1:].contiguous(), a[:, -1:].contiguous(), o[:, -1:].contiguous(), first_tran=ft, use_graph=False)   # one frame -> one launch per phase in the trace
torch.cuda.synchronize()
names = {0: 'P1 rnn4+rnn2', 2: 'P2 rnn6+rnn3+rnn7+rnn8', 3: 'LATE rnn4+rnn6'}
for ph in (0, 2, 3):
    buf = np.zeros((20000, 16), dtype=np.int64)
    nt = ctypes.c_int(0)
    _lib.check(lib.rc_state_debug_phase_trace(st, 0, ph, buf.ctypes.data_as(ctypes.c_void_p), buf.shape[0], ctypes.byref(nt)))
    tr = buf[:nt.value + 1]
    glob = tr[tr[:, 0] < 0]                       # kernel-wide stamps (complemented minima are negative as int64)
    tr = tr[(tr[:, 1] != 0) & (tr[:, 0] >= 0)]
    cta = tr[:, 0] >> 32
    job = (tr[:, 0] >> 16) & 0xffff
    print('==== phase %d (%s): %d tiles on %d CTAs' % (ph, names[ph], len(tr), len(np.unique(cta))))
    dep = tr[:, 2] - tr[:, 1]
    fill = tr[:, 3] - tr[:, 2]
    mma = tr[:, 4] - tr[:, 3]
    lag = tr[:, 5] - tr[:, 4]
    epa = tr[:, 6] - tr[:, 5]
    epb = tr[:, 7] - tr[:, 6]
    tot = tr[:, 7] - tr[:, 1]
    print(' job  tiles |  dep-wait   grab->mma0  mma-issue  commit->epi  epi math+st  publish   grab->done   (median cycles; max)')
    for jj in np.unique(job):
        s = job == jj
        f = lambda x: '%7d/%-8d' % (np.median(x[s]), x[s].max())
        print(' %3d %6d | %s %s %s %s %s %s %s' % (jj, s.sum(), f(dep), f(fill), f(mma), f(lag), f(epa), f(epb), f(tot)))
    pair_kernel = os.environ.get('RC_PH_PAIR', '0') not in ('', '0')        # slots 9-11 hold the MMA-thread waits only in the pair kernel
    for jj in (np.unique(job) if pair_kernel else []):
        s2 = job == jj
        print('   job %2d MMA thread waits (median cycles): tile descriptor %6d | accumulators handed back %6d | operand stages (sum over K blocks) %6d of %6d issue time' % (
            jj, np.median(tr[s2, 9]), np.median(tr[s2, 10]), np.median(tr[s2, 11]), np.median(mma[s2])))
    if os.environ.get('RC_PH_PAIR', '0') == '2':      # 256 x 256 pair kernel: epilogue breakdown (slots 5, 12, 8, 13, 14, 15, 6)
        for jj in np.unique(job):
            s2 = job == jj
            g = lambda a, b: int(np.median(tr[s2, a] - tr[s2, b]))
            print('   job %2d epilogue (median cycles): corr drain %6d | wait + main drain %6d | cprev load %6d | gate math %6d | C / H stores %6d | split stores %6d' % (
                jj, g(12, 5), g(8, 12), g(13, 8), g(14, 13), g(15, 14), g(6, 15)))
    if len(glob):
        gl = glob[0]
        t_entry, t_grab, t_pub, t_exit = ~int(gl[0]), ~int(gl[1]), int(gl[2]), int(gl[3])
        print(' kernel-wide (globaltimer): first CTA entry -> first tile grab %.1f us | first grab -> last publish %.1f us | last publish -> last CTA exit %.1f us' % (
            (t_grab - t_entry) * 1e-3, (t_pub - t_grab) * 1e-3, (t_exit - t_pub) * 1e-3))
    # per-CTA span and busy fraction
    spans, busy = [], []
    for c in np.unique(cta):
        s = cta == c
        spans.append(tr[s, 7].max() - tr[s, 1].min())
        busy.append(mma[s].sum())
    spans, busy = np.array(spans), np.array(busy)
    print(' per-CTA span: median %d max %d cycles; MMA-issue busy fraction median %.2f' % (np.median(spans), spans.max(), np.median(busy / spans)))
