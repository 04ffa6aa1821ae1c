import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import robustcap_b200 as rb
from robustcap_b200 import synthetic
B = int(sys.argv[1]); T = 300
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
net = rb.Net(rb.ParametricModel(assets['smpl_file']))
net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
inp = synthetic.make_inputs(B, T, seed=1000, conf='mixed')
rb.Net.gravityc = inp['gravity'].clone()
j, a, o = inp['j2dc'].cuda(), inp['accc'].cuda(), inp['oric'].cuda()
ft = torch.tensor([0., 0., 4.], device='cuda')
for mode in (2, 3):
    net.set_gemm_mode(mode)
    net.set_seq_options(auto_max_streams=0, warm_frames=16)
    for _ in range(2): net.forward_offline(j, a, o, first_tran=ft)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): net.forward_offline(j, a, o, first_tran=ft)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('B=%d mode %d BN=%s: %.1f us per frame, %.0f frames/s' % (B, mode, os.environ.get('RC_SEQ_BN', 'auto'), ms / T * 1e3, B * T / ms * 1e3), flush=True)
