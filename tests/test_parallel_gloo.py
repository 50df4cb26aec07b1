"""world_size-2 checks of the multi-GPU host logic on CPU (gloo): clip sharding,
max-over-ranks timing, caption gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from video_description_with_spatial_temporal_attention_b200 import parallel


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 257):
        for ws in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        lo, hi = parallel.shard_range(n_total, rank, ws)
        maxlen = 4
        toks = torch.arange(lo, hi)[:, None].repeat(1, maxlen) * 10 + torch.arange(maxlen)[None]
        lens = torch.arange(lo, hi, dtype=torch.int32) % maxlen + 1
        sc = torch.arange(lo, hi, dtype=torch.float32) * 0.5
        t, l, s = parallel.gather_captions(toks, lens, sc, n_total)
        mx = parallel.max_over_ranks([1.0 + rank, 5.0 - rank])
        q.put((rank, t.tolist(), l.tolist(), s.tolist(), mx))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_and_max_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    n_total = 7            # ragged shards: 4 + 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    want_t = [[i * 10 + j for j in range(4)] for i in range(n_total)]
    for rank, t, l, s, mx in res:
        assert t == want_t
        assert l == [i % 4 + 1 for i in range(n_total)]
        assert s == [i * 0.5 for i in range(n_total)]
        assert mx == [2.0, 5.0]


def _allreduce_worker(rank, world, port, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle import optim_oracle as oo
    from video_description_with_spatial_temporal_attention_b200 import optim
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    n = 1003
    local = np.random.RandomState(100 + rank).randn(n).astype('float32')
    g = torch.from_numpy(local.copy())
    optim.allreduce_grads(g)                              # SUM over the ranks, in place
    total = sum(np.random.RandomState(100 + r).randn(n).astype('float32') for r in range(world))
    ok = bool(np.allclose(g.numpy(), total, rtol=1e-6, atol=1e-6))
    # clipping happens after the all-reduce and the update is the same on every rank: identical parameters
    gc, _ = oo.clip(g.numpy(), 5.0)
    p = oo.Adam(n).update(np.zeros(n, 'float32'), gc)
    gathered = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(p))
    same = all(torch.equal(gathered[0], t) for t in gathered)
    q.put((rank, ok, same))
    dist.destroy_process_group()


def test_gradient_allreduce_then_clip_then_update_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok and same for _, ok, same in res), res
