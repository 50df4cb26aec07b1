"""Attention launch time when other traffic runs between two launches: nothing, the step's own GEMMs
(weights evict_first), or a plain 42 MB read (normal priority)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import synthetic
from video_description_with_spatial_temporal_attention_b200.engine import Engine
import bench

o = stat.baseline_options()
params = synthetic.trained_like_params(o, seed=7)
feats = bench.make_inputs(o, 1234)
eng = Engine(o)
eng.set_params(params)
devf = [torch.from_numpy(a).cuda() for a in feats]
ws, d = eng.precompute(*devf)
eng.decode_greedy(ws, d, 1)
torch.cuda.synchronize()
w = torch.randn(42 * 2**20 // 4, device='cuda')
A = torch.randn(64, 512, device='cuda')
WvT = torch.randn(12594, 512, device='cuda')     # logits-sized weight (25.8 MB)
WhT = torch.randn(4608, 512, device='cuda')      # h-projection-sized weight (9.4 MB)
WcT = torch.randn(2560, 512, device='cuda')      # ctx-projection-sized weight (5.2 MB)


def between(kind):
    if kind == 'gemms':
        eng.gemm(A, WcT, swap=True)
        eng.gemm(A, WhT, swap=True)
        eng.gemm(A, WvT, swap=True)
    elif kind == 'read42':
        w.sum()


for kind in ('nothing', 'gemms', 'read42', 'gemms'):
    for _ in range(5):
        eng.attention(ws, d)
        between(kind)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.attention(ws, d)
        b.record()
        between(kind)
        ts.append((a, b))
    torch.cuda.synchronize()
    print('%-8s attention %.2f us' % (kind, sum(a.elapsed_time(b) for a, b in ts) / len(ts) * 1e3))
