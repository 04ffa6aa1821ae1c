"""GPU aid (not a test): timing sweep of the grouped-kernel variants / tile policies (per-process environment switches), and the pose
difference of each configuration against the first one.  Usage: python tests/gpu_pair_sweep.py B T "K=V,K=V;K=V;..." [conf]"""
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gpu_pair_ab import angle

B, T = int(sys.argv[1]), int(sys.argv[2])
configs = sys.argv[3].split(';')
conf = sys.argv[4] if len(sys.argv) > 4 else 'mixed'
ref = None
for i, c in enumerate(configs):
    env = dict(os.environ)
    for kv in c.split(','):
        if kv:
            k, v = kv.split('=')
            env[k] = v
    out = '/tmp/sweep_%d.pt' % i
    print('--- [%s]' % c, flush=True)
    r = subprocess.run([sys.executable, os.path.join(HERE, 'gpu_pair_ab.py'), '--child', str(B), str(T), conf, out], env=env, timeout=600)
    if r.returncode != 0:
        print('FAILED rc=%d' % r.returncode, flush=True)
        continue
    res = torch.load(out)
    if ref is None:
        ref = res
    else:
        an = angle(res['pose'], ref['pose'])
        print('    vs first config: pose max %.2e rad, 99.9 %% %.2e, tran max %.2e m, finite %s' % (
            an.max().item(), an.flatten().kthvalue(int(an.numel() * 0.999)).values.item(), (res['tran'] - ref['tran']).abs().max().item(),
            bool(torch.isfinite(res['pose']).all())), flush=True)
        per_stream = an.amax(dim=(1, 2))
        top = per_stream.topk(min(5, len(per_stream)))
        print('    worst streams: ' + ', '.join('%d: %.2e' % (i, v) for v, i in zip(top.values.tolist(), top.indices.tolist())), flush=True)
