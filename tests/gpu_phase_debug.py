"""GPU debugging aid (not a test): persistent grouped tcgen05 kernel (gemm mode 2) against the per-layer tcgen05 path (mode 1)
and the fp32 SIMT path (mode 0) on the same batch, plus a quick timing.  Usage: python tests/gpu_phase_debug.py [B] [T]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robustcap_b200 as rb
from robustcap_b200 import synthetic, _lib


def pose_angle(a, b):
    r = a.reshape(-1, 3, 3).transpose(1, 2) @ b.reshape(-1, 3, 3)
    c = ((r[:, 0, 0] + r[:, 1, 1] + r[:, 2, 2]) - 1) / 2
    return torch.acos(c.clamp(-1, 1))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
    net = rb.Net(rb.ParametricModel(assets['smpl_file']))
    net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
    inp = synthetic.make_inputs(B, T, seed=31, conf='mixed')
    rb.Net.gravityc = inp['gravity'].clone()
    j, a, o = inp['j2dc'].cuda(), inp['accc'].cuda(), inp['oric'].cuda()
    ft = torch.tensor([0., 0., 4.])
    ff = torch.arange(B) % 5 == 0
    outs = {}
    for mode in (0, 1, 2):
        net.set_gemm_mode(mode)
        for use_graph in (False, True):
            p, t = net.forward_offline(j, a, o, first_tran=ft, first_frame=ff, first_tran_mask=~ff, use_graph=use_graph)
            torch.cuda.synchronize()
            outs[(mode, use_graph)] = (p.cpu(), t.cpu(), {k: v.clone() for k, v in net.debug_outputs(B).items()})
        print('mode', mode, 'done', flush=True)
    for key in ((1, False), (2, False), (2, True)):
        p0, t0, d0 = outs[(0, False)]
        p1, t1, d1 = outs[key]
        for k in d0:
            print('  mode %s sub-net %d: max |simt - x| = %.2e (scale %.2e)' % (key, k, (d0[k] - d1[k]).abs().max().item(), d0[k].abs().max().item()))
        ang = pose_angle(p0, p1)
        print('mode %s vs simt: pose max %.2e rad, 99.9%% %.2e, tran %.2e m' % (key, ang.max().item(), ang.quantile(0.999).item(), (t0 - t1).abs().max().item()), flush=True)
    for mode in (1, 2):
        net.set_gemm_mode(mode)
        for _ in range(2):
            net.forward_offline(j, a, o, first_tran=ft)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        net.forward_offline(j, a, o, first_tran=ft)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print('mode %d: %.1f us per frame step (B=%d), %.0f frames/s' % (mode, dt / T * 1e6, B, B * T / dt), flush=True)


if __name__ == '__main__':
    main()
