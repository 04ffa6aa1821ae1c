"""GPU profiling aid (not a test): one short batched forward_offline in a chosen gemm mode, for `ncu` launch lists.
Usage: python tests/gpu_phase_time.py [mode] [B] [T] [conf]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robustcap_b200 as rb
from robustcap_b200 import synthetic

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
T = int(sys.argv[3]) if len(sys.argv) > 3 else 6
conf = sys.argv[4] if len(sys.argv) > 4 else 'mixed'
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
net = rb.Net(rb.ParametricModel(assets['smpl_file']))
net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
inp = synthetic.make_inputs(B, T, seed=31, conf=conf)
rb.Net.gravityc = inp['gravity'].clone()
net.set_gemm_mode(mode)
p, t = net.forward_offline(inp['j2dc'].cuda(), inp['accc'].cuda(), inp['oric'].cuda(), first_tran=torch.tensor([0., 0., 4.]), use_graph=False)
torch.cuda.synchronize()
print('ok', p.shape, float(t.abs().max()))
