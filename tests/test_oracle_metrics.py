"""Pin oracle/metrics.py against the reference's cal_mpjpe (tests/golden/metrics.npz). CPU only."""
import torch

from oracle.kinematics import BodyOracle
from oracle.metrics import cal_mpjpe
from test_oracle_golden import load


def test_cal_mpjpe(golden_dir, assets):
    g = load(golden_dir, 'metrics.npz')
    body = BodyOracle(assets['smpl_file'])
    r3 = cal_mpjpe(body, g['j_regressor'], g['pose'], g['gt_pose'], cal_pampjpe=True)
    r2 = cal_mpjpe(body, g['j_regressor'], g['pose'], g['gt_pose'])
    assert (r3 - g['with_pa']).abs().max() < 1e-6 and (r2 - g['without_pa']).abs().max() < 1e-6
