"""GPU experiment (not a pytest): a few forward_online frames for an ncu launch list of the B=1 path."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robustcap_b200 as rb
from robustcap_b200 import synthetic
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
net = rb.Net(rb.ParametricModel(assets['smpl_file']))
net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
inp = synthetic.make_inputs(1, 40, seed=5, conf='high')
for t in range(40):
    net.forward_online(inp['j2dc'][0, t], inp['accc'][0, t], inp['oric'][0, t])
torch.cuda.synchronize()
