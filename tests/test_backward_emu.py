"""The backward pass (csrc/backward.cu: kernels + host orchestration of stat_grad_shared) executed on the CPU
through tests/emu (threads-as-CUDA-threads emulation, the GEMM replaced by a loop with the same contract)
and compared with the gradient oracle (torch fp64 autograd over the restated forward, oracle/grad_oracle.py).
The forward products the entry point consumes (context blocks, attention weights, hidden states) come from the
numpy oracle in float32, as the CUDA forward would leave them.  This checks the derivation, the indexing,
the reductions and the launch sequence; what it cannot check (nvcc code generation, the tensor-core GEMM on
these operand layouts) is covered by tests/test_gpu_parity.py::test_grad_shared_* on the GPU."""
import ctypes as C

import numpy as np
import pytest

import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import _lib, synthetic
from oracle import grad_oracle as go, stat_oracle as so
from tests.emu import build_emu


class FwdBlocks(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('ctxg0', 'pctxg', 'ctxm0', 'pctxm', 'ctxl0', 'pctxl', 'qctxl', 'h0c0')]


def load_emu():
    lib = C.CDLL(build_emu.build())
    lib.stat_grad_workspace_bytes.restype = C.c_size_t
    lib.stat_grad_workspace_bytes.argtypes = [C.POINTER(_lib.StatDims), C.c_int]
    lib.stat_grad_shared.restype = C.c_int
    lib.stat_grad_shared.argtypes = ([C.POINTER(_lib.StatDims), C.POINTER(_lib.StatParams), C.POINTER(FwdBlocks), C.c_int]
                                     + [C.c_void_p] * 14 + [C.c_float] * 3
                                     + [C.POINTER(_lib.StatParams), C.c_void_p, C.c_void_p])
    return lib


@pytest.fixture(scope='module')
def emu():
    return load_emu()


def _ptr(a):
    return None if a is None else a.ctypes.data


def run_emu(lib, o, params, batch, alpha_c, decay_c, dp=None, inv_batch=None, flat=False):
    x, mask, ctxg, mask_ctxg, ctxl, _ml, ctxm, _mm = batch
    L, B = x.shape
    T, R = ctxl.shape[1], ctxl.shape[2]
    H, E, V = o['dim'], o['dim_word'], o['n_words']
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    dp = dp or {}
    dpk = {k: (f32(v) if v is not None else None) for k, v in dp.items()}
    _, allv = so.forward_teacher(params, o, *batch, dtype=np.float32, return_all=True, **dpk)
    blk = allv['blk']
    P32 = {k: f32(v) for k, v in params.items()}
    if flat:
        # one contiguous buffer in init_params order, like optim.FlatParams: everything after the first scalar
        # parameter is only 4-byte aligned
        buf = np.zeros(sum(v.size for v in P32.values()) + 4, np.float32)
        off = 0
        for k, v in list(P32.items()):
            buf[off:off + v.size] = v.reshape(-1)
            P32[k] = buf[off:off + v.size].reshape(v.shape)
            off += v.size
        assert any(v.ctypes.data % 16 for v in P32.values())
    keep = dict(ctxg0=f32(blk['G']), pctxg=f32(blk['pG']), ctxm0=f32(blk['M']), pctxm=f32(blk['pM']),
                ctxl0=f32(blk['Lc']), pctxl=f32(blk['pL']), qctxl=f32(blk['Lc'] @ P32['decoder_Wclt_att']),
                h0c0=f32(np.concatenate([allv['h0'], allv['c0']], 1)))
    fb = FwdBlocks(**{k: _ptr(v) for k, v in keep.items()})
    al = f32(np.stack([s['alphaL'] for s in allv['steps']]))
    ag = f32(np.stack([s['alphaG'] for s in allv['steps']]))
    am = f32(np.stack([s['alphaM'] for s in allv['steps']]))
    alt = f32(np.stack([s['alphaLT'] for s in allv['steps']]))
    hall = f32(np.stack([s['h'] for s in allv['steps']]))
    d = _lib.StatDims(B, T, R, ctxg.shape[2], ctxm.shape[2], ctxl.shape[3], H, E, V, _lib.flags_of(o))
    sp, sg = _lib.StatParams(), _lib.StatParams()
    grads = {k: np.full_like(v, np.nan) for k, v in P32.items()}           # every element must be written
    for k in _lib.PARAM_FIELDS:
        setattr(sp, k, _ptr(P32[k]) if k in P32 else None)
        setattr(sg, k, _ptr(grads[k]) if k in grads else None)
    nbytes = lib.stat_grad_workspace_bytes(C.byref(d), L)
    assert nbytes > 0
    gws = np.full(nbytes // 4, np.nan, np.float32)                         # nothing may rely on a zeroed workspace
    x64 = np.ascontiguousarray(x, dtype=np.int64)
    args = [f32(mask), f32(ctxg), f32(mask_ctxg), f32(ctxl), f32(ctxm)]
    rc = lib.stat_grad_shared(C.byref(d), C.byref(sp), C.byref(fb), L, _ptr(x64), *[_ptr(a) for a in args],
                              _ptr(dpk.get('dp_gates')), _ptr(dpk.get('dp_h')), _ptr(dpk.get('dp_z')),
                              _ptr(al), _ptr(ag), _ptr(am), _ptr(alt), _ptr(hall),
                              inv_batch if inv_batch is not None else 1.0 / B, alpha_c, decay_c, C.byref(sg),
                              _ptr(gws), None)
    assert rc == 0
    return grads


def _case(global_proj, selector=True, ctx2out=True, prev2out=True):
    kw = dict(dim=8, dim_word=8, ctxl_dim=12, ctxm_dim=16, n_words=11, selector=selector, ctx2out=ctx2out,
              prev2out=prev2out)
    o = stat.default_options(ctxg_dim=12, global_proj=True, **kw) if global_proj else \
        stat.default_options(ctxg_dim=8, **kw)
    params = so.trained_like_params(o, seed=5)
    batch = synthetic.make_batch(o, B=3, T=4, R=2, L=5, seed=5, zero_tail=True)
    return o, params, batch


def _compare(grads, want, rtol=2e-5):
    assert list(grads.keys()) == list(want.keys())
    for k, w in want.items():
        g = grads[k]
        assert np.isfinite(g).all(), k
        scale = max(float(np.abs(w).max()), 1e-6)
        err = float(np.abs(g.astype('float64') - w).max())
        assert err <= rtol * scale + 1e-6, (k, err, scale)


@pytest.mark.parametrize('global_proj', [False, True])
def test_grads_match_oracle(emu, global_proj):
    o, params, batch = _case(global_proj)
    kw = dict(alpha_c=0.70602, decay_c=1e-4)
    grads = run_emu(emu, o, params, batch, **kw)
    _, want, _ = go.cost_and_grads(params, o, batch, **kw)
    _compare(grads, want)
    # and against the committed fixture of the same case (tests/golden/grad_toy.npz)
    import os
    from collections import OrderedDict
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'grad_toy.npz'))
    _compare(grads, OrderedDict((k, gold['gp%d/%s' % (int(global_proj), k)]) for k in want))


@pytest.mark.parametrize('fast', [False, True])
def test_grads_without_regularisers_and_options_off(emu, fast, monkeypatch):
    monkeypatch.setenv('STAT_BW_FAST', '1' if fast else '0')
    o, params, batch = _case(False, selector=False, ctx2out=False, prev2out=False)
    grads = run_emu(emu, o, params, batch, 0., 0., flat=fast)
    _, want, _ = go.cost_and_grads(params, o, batch)
    _compare(grads, want)


@pytest.mark.parametrize('fast', [False, True])
def test_grads_with_dropout_factors(emu, fast, monkeypatch):
    monkeypatch.setenv('STAT_BW_FAST', '1' if fast else '0')
    o, params, batch = _case(True)
    L, B = batch[0].shape
    H, E = o['dim'], o['dim_word']
    rng = np.random.RandomState(2)
    dp = dict(dp_gates=rng.binomial(1, 0.5, (L, B, 3 * H)).astype('float32'),
              dp_h=rng.binomial(1, 0.5, (L, B, H)).astype('float32'),
              dp_z=rng.binomial(1, 0.5, (L, B, E)).astype('float32'))
    grads = run_emu(emu, o, params, batch, 0.3, 1e-4, dp=dp, flat=fast)
    _, want, _ = go.cost_and_grads(params, o, batch, alpha_c=0.3, decay_c=1e-4, **dp)
    _compare(grads, want)


@pytest.mark.parametrize('fast', [False, True])
def test_grads_wider_than_a_block(emu, fast, monkeypatch):
    """(fast: the default variants -- k-split products summed from planes, owner-block embedding scatter, aligned
    copies of the weights that sit 4-byte aligned in a flat parameter buffer.)
    H = 160 > the 128 threads of the per-frame blocks (strided column loops, two columns per thread for some
    threads only), R = 3, a vocabulary that is not a multiple of 4, a single-frame tail."""
    kw = dict(dim=160, dim_word=12, ctxl_dim=20, ctxm_dim=12, n_words=37)
    o = stat.default_options(ctxg_dim=24, global_proj=True, **kw)
    params = so.trained_like_params(o, seed=9)
    batch = synthetic.make_batch(o, B=2, T=5, R=3, L=3, seed=9, zero_tail=True)
    monkeypatch.setenv('STAT_BW_FAST', '1' if fast else '0')
    grads = run_emu(emu, o, params, batch, 0.70602, 1e-4, flat=fast)
    _, want, _ = go.cost_and_grads(params, o, batch, alpha_c=0.70602, decay_c=1e-4)
    _compare(grads, want)


# ---- data parallelism: two ranks over gloo, each with half of the clips ------------------------
def _dp_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    from video_description_with_spatial_temporal_attention_b200 import optim, parallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        o, params, _ = _case(True)
        batch = synthetic.make_batch(o, B=4, T=4, R=2, L=5, seed=6, zero_tail=True)
        B = batch[0].shape[1]
        lo, hi = parallel.shard_range(B, rank, world)
        shard = tuple(a[:, lo:hi] if i < 2 else a[lo:hi] for i, a in enumerate(batch))
        # every rank: NLL scaled by 1 / B_global, coverage summed over its clips, decay on rank 0 only
        g = run_emu(load_emu(), o, params, shard, 0.70602, 1e-4 if rank == 0 else 0., inv_batch=1.0 / B)
        flat = torch.from_numpy(np.concatenate([v.reshape(-1) for v in g.values()]))
        optim.allreduce_grads(flat)                       # the product's collective: one SUM over the flat buffer
        q.put((rank, flat.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_data_parallel_gradients_equal_the_global_batch():
    import socket
    import torch.multiprocessing as mp
    build_emu.build()                                    # once, before the ranks load it
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert np.array_equal(res[0], res[1])                 # every rank holds the same reduced gradient
    o, params, _ = _case(True)
    batch = synthetic.make_batch(o, B=4, T=4, R=2, L=5, seed=6, zero_tail=True)
    _, want, _ = go.cost_and_grads(params, o, batch, alpha_c=0.70602, decay_c=1e-4)
    off = 0
    got = {}
    for k, w in want.items():
        got[k] = res[0][off:off + w.size].reshape(w.shape)
        off += w.size
    _compare(got, want)


@pytest.mark.parametrize('B,T,R,L', [(1, 1, 1, 1), (2, 1, 3, 2), (1, 3, 1, 4), (2, 2, 16, 2)])
def test_grads_edge_shapes(emu, B, T, R, L):
    """One clip, one frame, one region (soft-maxes over a single element: alpha = 1, zero score gradients), a single
    step (no recurrence), the maximum region count."""
    o, params, _ = _case(True)
    batch = synthetic.make_batch(o, B=B, T=T, R=R, L=L, seed=B + 10 * T + 100 * R)
    grads = run_emu(emu, o, params, batch, 0.70602, 1e-4)
    _, want, _ = go.cost_and_grads(params, o, batch, alpha_c=0.70602, decay_c=1e-4)
    _compare(grads, want)
