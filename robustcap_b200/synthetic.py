"""Seeded synthetic stand-ins for the licensed assets the hot path loads, and synthetic workloads.

The reference needs ``models/SMPL_male.pkl`` (``articulate/model.py:29-39``), ``data/dataset_work/gmm_08.pkl``
(``net/smplify/prior.py:113-116``) and trained weights (``net/sig_mp.py:850-857``).  None of them can be
redistributed, and BASELINE.json asks for random-init weights and synthetic inputs, so everything here is
generated deterministically from integer seeds with the same on-disk schema as the real files.  A user with
the real assets passes the real paths instead.

Nothing here is on the timed path.
"""
import math
import os
import pickle

import numpy as np
import torch

from .constants import NET_DIMS, NET_ORDER, INIT_NET_DIMS, SMPL_PARENT, NUM_VERTS

# Rough metric rest-pose skeleton of an adult (x left, y up, z forward), pelvis at index 0.
_REST_JOINTS = np.array([
    [0.000, -0.240, 0.030], [0.070, -0.330, 0.020], [-0.070, -0.330, 0.020], [0.000, -0.130, 0.000],
    [0.100, -0.710, 0.020], [-0.100, -0.710, 0.020], [0.000, 0.010, 0.020], [0.090, -1.110, -0.030],
    [-0.090, -1.110, -0.030], [0.000, 0.060, 0.030], [0.110, -1.170, 0.090], [-0.110, -1.170, 0.090],
    [0.000, 0.270, -0.010], [0.080, 0.180, 0.000], [-0.080, 0.180, 0.000], [0.000, 0.340, 0.040],
    [0.170, 0.220, -0.010], [-0.170, 0.220, -0.010], [0.430, 0.210, -0.040], [-0.430, 0.210, -0.040],
    [0.680, 0.220, -0.040], [-0.680, 0.220, -0.040], [0.770, 0.210, -0.050], [-0.770, 0.210, -0.050],
], dtype=np.float64)


def make_smpl_dict(seed=0, num_verts=NUM_VERTS):
    """Synthetic SMPL-schema dict: keys as read by ``articulate/model.py:29-39``."""
    import scipy.sparse as sp
    rs = np.random.RandomState(seed)
    J = _REST_JOINTS + rs.normal(0, 0.004, _REST_JOINTS.shape)
    home = rs.randint(0, 24, size=num_verts)
    v_template = J[home] + rs.normal(0, 0.045, (num_verts, 3))
    # <=4 influences per vertex like the real model, but every 16th vertex is dense over all 24 joints so
    # that dense skinning rows are exercised too.
    weights = np.zeros((num_verts, 24))
    parent = np.array([0] + SMPL_PARENT[1:])
    for k in range(4):
        j = home if k == 0 else (parent[home] if k == 1 else rs.randint(0, 24, size=num_verts))
        weights[np.arange(num_verts), j] += rs.uniform(0.05, 1.0, num_verts) * (0.5 ** k)
    dense = np.arange(0, num_verts, 16)
    weights[dense] += rs.uniform(0.0, 0.05, (dense.size, 24))
    weights /= weights.sum(axis=1, keepdims=True)
    # sparse joint regressor, rows sum to one
    rows, cols, vals = [], [], []
    for j in range(24):
        idx = np.where(home == j)[0][:32]
        w = rs.uniform(0.1, 1.0, idx.size)
        w /= w.sum()
        rows += [j] * idx.size
        cols += idx.tolist()
        vals += w.tolist()
    J_regressor = sp.csc_matrix((vals, (rows, cols)), shape=(24, num_verts))
    kintree = np.zeros((2, 24), dtype=np.uint32)
    kintree[0] = np.array([2 ** 32 - 1] + SMPL_PARENT[1:], dtype=np.uint32)
    kintree[1] = np.arange(24)
    return {
        'J_regressor': J_regressor,
        'weights': weights,
        'posedirs': rs.normal(0, 1e-3, (num_verts, 3, 207)).astype(np.float32),
        'shapedirs': rs.normal(0, 1e-2, (num_verts, 3, 10)),
        'v_template': v_template,
        'J': J,
        'f': rs.randint(0, num_verts, size=(13776, 3)).astype(np.uint32),
        'kintree_table': kintree,
    }


def make_gmm_dict(seed=1, num_gaussians=8, dim=69):
    """Synthetic GMM pose prior: keys as read by ``net/smplify/prior.py:113-116`` (SPD covariances)."""
    rs = np.random.RandomState(seed)
    means = rs.normal(0, 0.15, (num_gaussians, dim))
    covars = np.zeros((num_gaussians, dim, dim))
    for m in range(num_gaussians):
        a = rs.normal(0, 1.0, (dim, dim)) / math.sqrt(dim)
        covars[m] = 0.04 * (a @ a.T) + 0.02 * np.eye(dim)
    w = rs.uniform(0.5, 1.5, num_gaussians)
    return {'means': means, 'covars': covars, 'weights': w / w.sum()}


def write_assets(root, seed=0):
    """Write the synthetic asset tree (relative paths of ``config.py:1-24``) under ``root``; idempotent."""
    os.makedirs(os.path.join(root, 'models'), exist_ok=True)
    os.makedirs(os.path.join(root, 'data', 'dataset_work'), exist_ok=True)
    smpl = os.path.join(root, 'models', 'SMPL_male.pkl')
    gmm = os.path.join(root, 'data', 'dataset_work', 'gmm_08.pkl')
    jr = os.path.join(root, 'data', 'dataset_work', 'J_regressor_h36m.npy')
    tag = '.tmp%d' % os.getpid()       # several ranks may write concurrently: private temp file, atomic rename
    if not os.path.exists(smpl):
        with open(smpl + tag, 'wb') as f:
            pickle.dump(make_smpl_dict(seed), f, protocol=2)
        os.replace(smpl + tag, smpl)
    if not os.path.exists(gmm):
        with open(gmm + tag, 'wb') as f:
            pickle.dump(make_gmm_dict(seed + 1), f, protocol=2)
        os.replace(gmm + tag, gmm)
    if not os.path.exists(jr):
        rs = np.random.RandomState(seed + 2)
        r = rs.uniform(0, 1, (17, NUM_VERTS)) * (rs.uniform(0, 1, (17, NUM_VERTS)) < 0.004)
        r /= r.sum(axis=1, keepdims=True)
        with open(jr + tag, 'wb') as f:
            np.save(f, r.astype(np.float32))
        os.replace(jr + tag, jr)
    return {'smpl_file': smpl, 'gmm_dir': os.path.join(root, 'data', 'dataset_work'), 'j_regressor': jr}


def default_asset_root():
    return os.environ.get('ROBUSTCAP_ASSETS', os.path.join('/tmp', 'robustcap_b200_assets'))


def state_dict_keys():
    """The key set ``Net.load_state_dict`` accepts, in a fixed order, with shapes."""
    keys = []
    for name in NET_ORDER:
        i, h, o = NET_DIMS[name]
        for l in range(2):
            keys += [(f'{name}.rnn.weight_ih_l{l}', (4 * h, h)), (f'{name}.rnn.weight_hh_l{l}', (4 * h, h)),
                     (f'{name}.rnn.bias_ih_l{l}', (4 * h,)), (f'{name}.rnn.bias_hh_l{l}', (4 * h,))]
        keys += [(f'{name}.linear1.weight', (h, i)), (f'{name}.linear1.bias', (h,)),
                 (f'{name}.linear2.weight', (o, h)), (f'{name}.linear2.bias', (o,))]
        if name == 'rnn2':
            d = INIT_NET_DIMS
            for n, k in enumerate((0, 2, 4)):
                keys += [(f'rnn2.init_net.{k}.weight', (d[n + 1], d[n])), (f'rnn2.init_net.{k}.bias', (d[n + 1],))]
    return keys


def make_state_dict(seed=0, variant='default'):
    """Random-init weights with torch's default distributions (U(-1/sqrt(fan), 1/sqrt(fan))), generated
    tensor by tensor from one seeded CPU generator so the result does not depend on module construction
    order.  ``variant`` perturbs a few output biases so that data-dependent branches of
    ``net/sig_mp.py:185-225`` (foot contact, drift snap, floor) are reachable with untrained weights:

    * ``'default'``  – plain random init (contact ~0.5 -> velocity branch only).
    * ``'contact'``  – rnn8 output bias +1.6/+1.2 and a larger output gain -> contact branch + floor logic.
    * ``'snap'``     – as ``'contact'`` plus rnn6 output bias (0, 0, 14) -> ``|pc - tran| > 10`` snap on early frames.
    """
    g = torch.Generator(device='cpu')
    g.manual_seed(int(seed))
    sd = {}
    for key, shape in state_dict_keys():
        if '.rnn.' in key:
            bound = 1.0 / math.sqrt(shape[0] // 4)
        elif key.endswith('weight'):
            bound = 1.0 / math.sqrt(shape[1])
        else:
            w = sd[key[:-4] + 'weight']
            bound = 1.0 / math.sqrt(w.shape[1])
        sd[key] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    if variant in ('contact', 'snap'):
        sd['rnn8.linear2.weight'] = sd['rnn8.linear2.weight'] * 40.0
        sd['rnn8.linear2.bias'] = torch.tensor([1.6, 1.2])
    if variant == 'snap':
        sd['rnn6.linear2.bias'] = torch.tensor([0.0, 0.0, 14.0])
    elif variant != 'default' and variant != 'contact':
        raise ValueError(variant)
    return sd


def _random_rotations(n, g):
    """Uniform random rotations from normalised Gaussian quaternions (wxyz)."""
    q = torch.randn(n, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    a, b, c, d = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    r = torch.stack((1 - 2 * c * c - 2 * d * d, 2 * b * c - 2 * a * d, 2 * a * c + 2 * b * d,
                     2 * b * c + 2 * a * d, 1 - 2 * b * b - 2 * d * d, 2 * c * d - 2 * a * b,
                     2 * b * d - 2 * a * c, 2 * a * b + 2 * c * d, 1 - 2 * b * b - 2 * c * c), dim=1)
    return r.view(n, 3, 3)


def make_inputs(B, T, seed=0, conf='mixed', device='cpu'):
    """Synthetic workload of SURVEY.md §8(d): returns dict of float32 tensors
    ``j2dc[B,T,33,3]`` (x, y on the z=1 plane, confidence), ``accc[B,T,6,3]``, ``oric[B,T,6,3,3]``,
    ``gravity[3]``.

    ``conf``: ``'mixed'`` U(0.6, 1.0) per frame (all three branches), ``'high'`` U(0.9, 1),
    ``'mid'`` U(0.71, 0.79), ``'low'`` U(0.2, 0.69), ``'occluded'`` conf == 0 on a seeded random 50 % of
    frames and U(0.9, 1) elsewhere (BASELINE.json configs[4]).
    """
    g = torch.Generator(device='cpu')
    g.manual_seed(int(seed) + 7919)
    oric = _random_rotations(B * T * 6, g).view(B, T, 6, 3, 3)
    accc = torch.randn(B, T, 6, 3, generator=g)
    xy = torch.randn(B, T, 33, 2, generator=g) * 0.2
    u = torch.rand(B, T, 1, generator=g)
    if conf == 'mixed':
        c = 0.6 + 0.4 * u
    elif conf == 'high':
        c = 0.9 + 0.1 * u
    elif conf == 'mid':
        c = 0.71 + 0.08 * u
    elif conf == 'low':
        c = 0.2 + 0.49 * u
    elif conf == 'occluded':
        occ = torch.rand(B, T, 1, generator=g) < 0.5
        c = torch.where(occ, torch.zeros_like(u), 0.9 + 0.1 * u)
    else:
        raise ValueError(conf)
    j2dc = torch.cat((xy, c.unsqueeze(-1).expand(B, T, 33, 1)), dim=-1).contiguous()
    grav = torch.tensor([0.05, -0.99, 0.12])
    grav = grav / grav.norm()
    out = {'j2dc': j2dc, 'accc': accc, 'oric': oric, 'gravity': grav}
    return {k: v.to(device) for k, v in out.items()}
