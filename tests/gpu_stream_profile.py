"""GPU aid (not a test): B = 1 streaming latency (forward_online p50, device time per frame) for the current single-stream path.
RC_STREAM2=0 selects the multi-launch CUDA-graph path, RC_STREAM_TS=1 prints the in-kernel phase times."""
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robustcap_b200 as rb
from robustcap_b200 import synthetic

conf = sys.argv[1] if len(sys.argv) > 1 else 'high'
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
net = rb.Net(rb.ParametricModel(assets['smpl_file']))
net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
s1 = synthetic.make_inputs(1, 300, seed=5, conf=conf)
rb.Net.gravityc = s1['gravity'].clone()
net.reset_states()
lat, out = [], []
for t in range(300):
    t0 = time.perf_counter()
    p, tr = net.forward_online(s1['j2dc'][0, t], s1['accc'][0, t], s1['oric'][0, t])
    lat.append(time.perf_counter() - t0)
    out.append((p, tr))
net.reset_states()
dj, da, do = s1['j2dc'][0].cuda(), s1['accc'][0].cuda(), s1['oric'][0].cuda()
for _ in range(2):
    po, to = net.forward_offline(dj, da, do)
torch.cuda.synchronize()
g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g0.record()
po, to = net.forward_offline(dj, da, do)
g1.record()
torch.cuda.synchronize()
dev_us = 1e3 * g0.elapsed_time(g1) / 300
d = max((po[t].cpu() - out[t][0]).abs().max().item() for t in range(300))
print('conf=%s: forward_online p50 %.1f us, device %.1f us per frame (%.0f GB/s of weights), online vs offline max |dpose| %.2e, finite %s'
      % (conf, 1e6 * statistics.median(lat[20:]), dev_us, 243.06e6 / (dev_us * 1e-6) / 1e9, d, bool(torch.isfinite(po).all())))
torch.save({'pose': po.cpu(), 'tran': to.cpu()}, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gpurun_out', 'stream_%s_%s.pt' % (conf, os.environ.get('RC_STREAM2', '1'))))
