"""CPU: the oracle of SURVEY 8(f) rows 2 / 3 (oracle/pipeline.py) against outputs of the unmodified reference (golden)."""
import os

import numpy as np
import torch

from oracle import pipeline as OP
from oracle.kinematics import BodyOracle


def load(golden_dir):
    z = np.load(os.path.join(golden_dir, 'pipeline.npz'))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_pack_rows_oracle_vs_reference(golden_dir):
    g = load(golden_dir)
    for r, (i, j) in enumerate(g['rows'].tolist()):
        o = OP.pack_row(g['row%d_j2d' % r], g['seq%d_acc' % i], g['seq%d_ori' % i], g['row%d_cam_T' % r], g['row%d_cam_K' % r],
                        pose_aa=g['seq%d_pose_aa' % i], tran=g['seq%d_tran' % i])
        for k in ('j2dc', 'accc', 'oric', 'gravity', 'tran'):
            assert torch.equal(o[k], g['row%d_%s' % (r, k)]), (r, k)          # same torch ops in the same order: bit-exact
        assert (o['pose'] - g['row%d_pose' % r]).abs().max().item() < 5e-7    # the oracle's own Rodrigues restatement


def test_syn_acc_oracle_vs_reference(golden_dir):
    g = load(golden_dir)
    v = g['imu_mean_vimu']
    assert torch.equal(OP.syn_acc(v), g['imu_mean_acc'])
    assert torch.equal(OP.syn_acc(v, 4), g['imu_mean_acc4'])
    assert torch.equal(OP.syn_acc(v[:5]), g['imu_short_acc'])


def test_synthesize_imu_oracle_vs_reference(golden_dir, assets):
    g = load(golden_dir)
    body = BodyOracle(assets['smpl_file'])
    for tag, shape in (('mean', None), ('shaped', g['imu_shape'])):
        acc, ori, joint, vimu = OP.synthesize_imu(body, g['imu_pose'], g['imu_tran'], shape)
        assert (vimu - g['imu_%s_vimu' % tag]).abs().max().item() < 2e-6, tag
        assert (joint - g['imu_%s_joint' % tag]).abs().max().item() < 2e-6, tag
        assert (ori - g['imu_%s_ori' % tag]).abs().max().item() < 2e-6, tag
        # second differences x 3600 amplify 1e-6 m of float32 FK noise to ~1e-2 m/s^2
        assert (acc - g['imu_%s_acc' % tag]).abs().max().item() < 3e-2, tag
        assert torch.equal(OP.syn_acc(g['imu_%s_vimu' % tag]), g['imu_%s_acc' % tag])
