r"""Multi-GPU offline evaluation: sequences are independent (all state is per stream, net/sig_mp.py:85-90), so a batch is
sharded by sequence across ranks with NO data-path collective; the single exchange is the final gather of the results
(``pose[B,T,24,3,3]``, ``tran[B,T,3]``) to rank 0 over NCCL/NVLink.  One process per GPU.
"""
import torch
import torch.distributed as dist

__all__ = ['shard_bounds', 'forward_offline_sharded', 'gather_results']


def shard_bounds(n, rank, world):
    """Contiguous block of sequences owned by ``rank``: sizes differ by at most one, order preserved."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local, n_total, dst=0, group=None):
    """Gather ragged per-rank blocks ``local`` [n_r, ...] (same trailing shape) to ``dst``; returns the concatenated tensor
    [n_total, ...] on ``dst`` and None elsewhere.  Uses NCCL send/recv of exact-size blocks when counts differ (gather needs
    equal sizes), plain gather otherwise."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    if len(set(counts)) == 1:
        out = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
        dist.gather(local.contiguous(), out, dst=dst, group=group)
        return torch.cat(out, dim=0) if rank == dst else None
    if rank == dst:
        parts = []
        for r in range(world):
            if r == dst:
                parts.append(local)
            else:
                buf = torch.empty((counts[r],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
                dist.recv(buf, src=r, group=group)
                parts.append(buf)
        return torch.cat(parts, dim=0)
    dist.send(local.contiguous(), dst=dst, group=group)
    return None


def forward_offline_sharded(forward, j2dc, accc, oric, dst=0, group=None, **kwargs):
    """Run ``forward(j2dc_shard, accc_shard, oric_shard, **kwargs_shard) -> (pose, tran)`` (e.g. ``Net.forward_offline``) on this
    rank's contiguous shard of the ``[B, T, ...]`` batch and gather the results to ``dst``.  Per-sequence keyword tensors
    (``first_tran`` [B,3], ``first_frame`` bool[B], ``lengths`` int[B], ``first_tran_mask`` bool[B]) are sharded the same way."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = j2dc.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    kw = {}
    for k, v in kwargs.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B:
            kw[k] = v[lo:hi]
        else:
            kw[k] = v
    pose, tran = forward(j2dc[lo:hi], accc[lo:hi], oric[lo:hi], **kw)
    return gather_results(pose, B, dst, group), gather_results(tran, B, dst, group)
