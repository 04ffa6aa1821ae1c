"""GPU profiling aid (not a test): one launch of the device-resident L-BFGS (rc_smplify_run) on 8 sequences x 100 frames."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robustcap_b200 as rb
from robustcap_b200 import synthetic, smplify

assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
os.chdir(os.path.dirname(os.path.dirname(assets['gmm_dir'])))
body = rb.ParametricModel(assets['smpl_file'])
S, T = 8, 100
g = torch.Generator().manual_seed(1)
pose = synthetic._random_rotations(S * T * 24, g).view(S, T, 24, 3, 3)
pose = torch.matrix_exp(0.2 * (torch.randn(S, T, 24, 3, 3, generator=g) - torch.randn(S, T, 24, 3, 3, generator=g).transpose(-1, -2)))
tran = torch.tensor([0., 0., 4.]) + 0.05 * torch.randn(S, T, 3, generator=g)
cam_k = torch.tensor([[1000.0, 0, 960], [0, 1000, 540], [0, 0, 1]])
_, kp3 = body.keypoints33(pose.reshape(-1, 24, 3, 3).cuda(), tran.reshape(-1, 3).cuda())
uv = (cam_k.cuda() @ (kp3 / kp3[..., 2:]).unsqueeze(-1)).squeeze(-1)[..., :2].reshape(S, T, 33, 2).cpu() + 5 * torch.randn(S, T, 33, 2, generator=g)
kp = torch.cat((uv, 0.5 + 0.5 * torch.rand(S, T, 33, 1, generator=g)), dim=-1)
ori = synthetic._random_rotations(S * T * 6, g).view(S, T, 6, 3, 3)
p2, t2, stats = smplify.smplify_runner_batch(pose, tran, kp, ori, cam_k, lr=1e-3, max_iter=5, body_model=body)
torch.cuda.synchronize()
print('stats', stats.cpu().tolist())
