r"""Multi-GPU offline evaluation: sequences are independent (all state is per stream, net/sig_mp.py:85-90), so a batch is
sharded by sequence across ranks with NO data-path collective; the single exchange is the final gather of the results
(``pose[B,T,24,3,3]``, ``tran[B,T,3]``) to rank 0 over NCCL/NVLink.  One process per GPU.

The results of a rank live in ONE flat buffer ``[pose | tran]`` (the kernels write straight into views of it), so the
exchange is a single collective (``gather`` when every rank owns the same number of sequences, exact-size send/recv
otherwise) instead of one per tensor.
"""
import torch
import torch.distributed as dist

__all__ = ['shard_bounds', 'forward_offline_sharded', 'gather_results', 'ShardedResult']

# keyword arguments of Net.forward_offline that carry one entry per sequence (everything else is passed through unchanged)
PER_SEQUENCE = ('first_tran', 'first_frame', 'lengths', 'first_tran_mask', 'gravity')


def shard_bounds(n, rank, world):
    """Contiguous block of sequences owned by ``rank``: sizes differ by at most one, order preserved."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _counts(n_total, world):
    return [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]


def gather_results(local, n_total, dst=0, group=None):
    """Gather ragged per-rank blocks ``local`` [n_r, ...] (same trailing shape; n_r may be 0) to ``dst``; returns the
    concatenated tensor [n_total, ...] on ``dst`` and None elsewhere.  Exact-size send/recv when counts differ (gather needs
    equal sizes), plain gather otherwise; ranks that own nothing neither send nor are received from."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = _counts(n_total, world)
    if len(set(counts)) == 1:
        out = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
        dist.gather(local.contiguous(), out, dst=dst, group=group)
        return torch.cat(out, dim=0) if rank == dst else None
    if rank == dst:
        parts = []
        for r in range(world):
            if r == dst:
                parts.append(local)
            elif counts[r] > 0:
                buf = torch.empty((counts[r],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
                dist.recv(buf, src=r, group=group)
                parts.append(buf)
        return torch.cat(parts, dim=0)
    if counts[rank] > 0:
        dist.send(local.contiguous(), dst=dst, group=group)
    return None


class ShardedResult:
    """Flat per-rank result buffer ``[pose (n*T*216) | tran (n*T*3)]`` with views for the kernels to write into."""

    def __init__(self, n, T, device, dtype=torch.float32):
        self.n, self.T = n, T
        self.flat = torch.zeros(n * T * 219, device=device, dtype=dtype)
        self.pose = self.flat[:n * T * 216].view(n, T, 24, 3, 3)
        self.tran = self.flat[n * T * 216:].view(n, T, 3)

    @staticmethod
    def split(flat, n, T):
        return flat[:n * T * 216].view(n, T, 24, 3, 3), flat[n * T * 216:].view(n, T, 3)


def gather_flat(res, n_total, dst=0, group=None, recv_bufs=None):
    """One exchange for pose AND tran: every rank contributes ``res.flat``; ``dst`` gets per-rank views (no concatenation copy
    when ``recv_bufs`` — a list of preallocated flat buffers, one per rank — is given).  Returns (pose, tran) lists on ``dst``."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = _counts(n_total, world)
    T = res.T
    if rank == dst and recv_bufs is None:
        recv_bufs = [torch.empty(c * T * 219, device=res.flat.device, dtype=res.flat.dtype) for c in counts]
    if len(set(counts)) == 1:
        dist.gather(res.flat, recv_bufs if rank == dst else None, dst=dst, group=group)
    elif rank == dst:
        for r in range(world):
            if r == dst:
                recv_bufs[r].copy_(res.flat)
            elif counts[r] > 0:
                dist.recv(recv_bufs[r], src=r, group=group)
    elif counts[rank] > 0:
        dist.send(res.flat, dst=dst, group=group)
    if rank != dst:
        return None, None
    parts = [ShardedResult.split(recv_bufs[r], counts[r], T) for r in range(world)]
    return [p for p, _ in parts], [t for _, t in parts]


def forward_offline_sharded(forward, j2dc, accc, oric, dst=0, group=None, fused=False, recv_bufs=None, **kwargs):
    """Run ``forward(j2dc_shard, accc_shard, oric_shard, **kwargs_shard) -> (pose, tran)`` (e.g. ``Net.forward_offline``) on this
    rank's contiguous shard of the ``[B, T, ...]`` batch and gather the results to ``dst``.

    Only the keyword arguments named in ``PER_SEQUENCE`` are sharded, and only when they really carry one entry per sequence
    (``first_tran`` must be ``[B, 3]`` — a shared ``[3]`` vector is passed through even when ``B == 3``).  A rank whose shard is
    empty (``B < world``) skips ``forward`` and takes part in the exchange with a zero-row block.
    ``fused=True``: ``forward`` must accept ``out=(pose, tran)``; results are written into one flat buffer and exchanged with a
    single collective; returns lists of per-rank views on ``dst`` (concatenate only if a single tensor is needed)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, T = j2dc.shape[0], j2dc.shape[1]
    lo, hi = shard_bounds(B, rank, world)
    kw = {}
    for k, v in kwargs.items():
        per_seq = k in PER_SEQUENCE and torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B
        if k == 'first_tran' and torch.is_tensor(v) and v.dim() != 2:
            per_seq = False
        kw[k] = v[lo:hi] if per_seq else v
    if fused:
        res = ShardedResult(hi - lo, T, j2dc.device)
        if hi > lo:
            forward(j2dc[lo:hi], accc[lo:hi], oric[lo:hi], out=(res.pose, res.tran), **kw)
        return gather_flat(res, B, dst, group, recv_bufs)
    if hi > lo:
        pose, tran = forward(j2dc[lo:hi], accc[lo:hi], oric[lo:hi], **kw)
    else:
        pose = torch.zeros(0, T, 24, 3, 3, device=j2dc.device)
        tran = torch.zeros(0, T, 3, device=j2dc.device)
    return gather_results(pose, B, dst, group), gather_results(tran, B, dst, group)
