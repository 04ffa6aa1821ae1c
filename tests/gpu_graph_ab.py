"""GPU aid (not a test): A/B timing of forward_offline with and without CUDA-graph replay, same process, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robustcap_b200 as rb
from robustcap_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
net = rb.Net(rb.ParametricModel(assets['smpl_file']))
net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
inp = synthetic.make_inputs(B, T, seed=1000, conf='mixed')
rb.Net.gravityc = inp['gravity'].clone()
j, a, o = inp['j2dc'].cuda(), inp['accc'].cuda(), inp['oric'].cuda()
ft = torch.tensor([0., 0., 4.], device='cuda')
for rep in range(2):
    for g in (False, True):
        for _ in range(2):
            net.forward_offline(j, a, o, first_tran=ft, use_graph=g)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            net.forward_offline(j, a, o, first_tran=ft, use_graph=g)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print('use_graph=%s: %.1f ms per pass, %.1f us per frame step, %.0f frames/s' % (g, ms, ms / T * 1e3, B * T / ms * 1e3), flush=True)
