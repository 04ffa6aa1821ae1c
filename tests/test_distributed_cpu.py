"""world_size-2 gloo test (CPU) of the sequence sharding + final gather used for multi-GPU offline evaluation.
The per-shard compute is the CPU oracle here (the test runs without a GPU); on the GPU box the same plumbing wraps
``Net.forward_offline`` with NCCL."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, ret):
    sys.path.insert(0, REPO)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from robustcap_b200 import synthetic
    from robustcap_b200.distributed import forward_offline_sharded, shard_bounds
    from oracle.kinematics import BodyOracle
    from oracle.fusion import FusionOracle
    assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
    sd = synthetic.make_state_dict(0, 'default')
    inp = synthetic.make_inputs(B, 4, seed=3, conf='mixed')
    o = FusionOracle(sd, BodyOracle(assets['smpl_file']))

    def forward(j, a, r, first_tran=None, out=None):
        ps, ts = [], []
        for b in range(j.shape[0]):
            ftb = None if first_tran is None else (first_tran[b] if first_tran.dim() == 2 else first_tran)
            p, t = o.run(j[b], a[b], r[b], first_tran=ftb, gravity=inp['gravity'])
            ps.append(p)
            ts.append(t)
        ps, ts = torch.stack(ps), torch.stack(ts)
        if out is not None:
            out[0].copy_(ps)
            out[1].copy_(ts)
            return out
        return ps, ts

    ft = torch.arange(B * 3, dtype=torch.float32).view(B, 3)
    pose, tran = forward_offline_sharded(forward, inp['j2dc'], inp['accc'], inp['oric'], dst=0, first_tran=ft)
    # a shared [3] first_tran must NOT be sliced, also when B == 3 (it would look like a per-sequence tensor)
    shared = torch.tensor([0., 0., 4.])
    pose_s, tran_s = forward_offline_sharded(forward, inp['j2dc'], inp['accc'], inp['oric'], dst=0, first_tran=shared)
    # one flat buffer, one exchange
    pl, tl = forward_offline_sharded(forward, inp['j2dc'], inp['accc'], inp['oric'], dst=0, fused=True, first_tran=ft)
    if rank == 0:
        full_p, full_t = forward(inp['j2dc'], inp['accc'], inp['oric'], first_tran=ft)
        ok = torch.equal(pose, full_p) and torch.equal(tran, full_t) and pose.shape[0] == B
        sh_p, sh_t = forward(inp['j2dc'], inp['accc'], inp['oric'], first_tran=shared)
        ok = ok and torch.equal(pose_s, sh_p) and torch.equal(tran_s, sh_t)
        ok = ok and torch.equal(torch.cat(pl), full_p) and torch.equal(torch.cat(tl), full_t)
        ret['ok'] = bool(ok)
        ret['bounds'] = [shard_bounds(B, r, world) for r in range(world)]
    else:
        assert pose is None and tran is None and pl is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [4, 3, 1])      # 1: fewer sequences than ranks -> one rank owns an empty shard
def test_sharded_gather_gloo(B):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000 + B
    mp.spawn(_worker, args=(2, port, B, ret), nprocs=2, join=True)
    assert ret['ok']
    assert ret['bounds'][0][0] == 0 and ret['bounds'][-1][1] == B
