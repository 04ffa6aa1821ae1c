"""GPU aid (not a test): where do the sequence kernel and the grouped kernel differ?  python tests/gpu_seq_debug.py B conf ragged"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import robustcap_b200 as rb
from robustcap_b200 import synthetic
from test_oracle_golden import pose_angle
B = int(sys.argv[1]); conf = sys.argv[2]; ragged = int(sys.argv[3])
T = 72
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
net = rb.Net(rb.ParametricModel(assets['smpl_file']))
net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
inp = synthetic.make_inputs(B, T, seed=900 + B, conf=conf)
rb.Net.gravityc = inp['gravity'].clone()
ff = torch.arange(B) % 3 == 0
lengths = (torch.arange(B) * 7 % (T - 20) + 20).to(torch.int32) if ragged else None
j, a, o = inp['j2dc'].cuda(), inp['accc'].cuda(), inp['oric'].cuda()
kw = dict(first_tran=torch.tensor([0., 0., 4.]), first_frame=ff, first_tran_mask=~ff, lengths=lengths)
out = {}
for mode in (2, 3):
    net.set_gemm_mode(mode)
    net.set_seq_options(auto_max_streams=0, warm_frames=16)
    p, t = net.forward_offline(j, a, o, **kw)
    out[mode] = (p.cpu(), t.cpu())
d = (out[2][0] - out[3][0]).abs().amax(dim=(2, 3, 4))
print('max diff', d.max().item(), 'rows differing', (d.amax(dim=1) > 0).nonzero().flatten().tolist()[:40])
bad = (d > 0).nonzero()
print('first (row, frame):', bad[:10].tolist())
if lengths is not None and len(bad):
    for r, t in bad[:10].tolist():
        print(r, t, 'length', int(lengths[r]), 'diff', d[r, t].item(), 'conf', inp['j2dc'][r, t, 0, 2].item())
# narrow down: rerun with T = first differing frame + 1 and compare the sub-net outputs of that frame
if len(bad):
    r0, t0 = bad[0].tolist()
    Tn = t0 + 1
    outs = {}
    for mode in (2, 3):
        net.set_gemm_mode(mode)
        net.set_seq_options(auto_max_streams=0, warm_frames=16)
        kw2 = dict(kw)
        if lengths is not None:
            kw2['lengths'] = torch.clamp(lengths, max=Tn)
        net.forward_offline(j[:, :Tn].contiguous(), a[:, :Tn].contiguous(), o[:, :Tn].contiguous(), **kw2)
        torch.cuda.synchronize()
        outs[mode] = net.debug_outputs(B)
        import ctypes
        from robustcap_b200 import _lib
        tb = torch.empty(B, 69)
        _lib.check(_lib.load().rc_state_debug_output(net._states[B], 70 if mode == 2 else 71, _lib.hptr(tb), _lib.stream()))
        outs[mode]['j3dr'] = tb
    print('blend row', r0, 'mode2', outs[2]['j3dr'][r0, :6].tolist(), 'mode3', outs[3]['j3dr'][r0, :6].tolist())
    dj = (outs[2]['j3dr'] - outs[3]['j3dr']).abs()
    print('blend diff max', dj.max().item(), 'rows', (dj.amax(dim=1) > 0).nonzero().flatten().tolist()[:10], 'cols', (dj[r0] > 0).nonzero().flatten().tolist())
    for k in (3, 6, 7, 8):
        dd = (outs[2][k] - outs[3][k]).abs().amax(dim=1)
        print('net', k, 'max diff', dd.max().item(), 'rows', (dd > 0).nonzero().flatten().tolist()[:10])
    confs = inp['j2dc'][r0, max(0, t0 - 6):t0 + 1, 0, 2].tolist()
    print('row', r0, 'conf history', ['%.3f' % c for c in confs])
    print('rows ending near:', [(b, int(lengths[b])) for b in range(B) if lengths is not None and abs(int(lengths[b]) - t0) <= 1])
