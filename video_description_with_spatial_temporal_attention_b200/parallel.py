"""Multi-GPU plumbing of the decode path: clips are independent, so
ranks take disjoint slices of the clip list and nothing crosses ranks on the data
path.  The only communication is for reporting: the max-over-ranks of a device
time and the gather of the finished captions.  One process per GPU,
torch.distributed (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank, world_size):
    """[lo, hi) of the items rank owns: contiguous, sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(values, device=None):
    """Element-wise max over ranks of a list of floats (device timings)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def gather_captions(tokens, lengths, scores, n_total):
    """Every rank passes the captions of its shard (torch tensors, shard order);
    returns the (n_total, maxlen) / (n_total,) / (n_total,) tensors in clip order
    on every rank.  Shards may have different sizes."""
    rank, ws = world()
    if ws == 1:
        return tokens, lengths, scores
    maxlen = tokens.shape[1]
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    cap = max(hi - lo for lo, hi in sizes)

    def pad(t, fill):
        out = torch.full((cap,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
        out[:t.shape[0]] = t
        return out

    outs = []
    for t, fill in ((tokens, -1), (lengths, 0), (scores, 0)):
        p = pad(t, fill)
        bufs = [torch.empty_like(p) for _ in range(ws)]
        dist.all_gather(bufs, p)
        outs.append(torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0))
    assert outs[0].shape == (n_total, maxlen)
    return tuple(outs)
