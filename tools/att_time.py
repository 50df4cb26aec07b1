"""att_group_kernel alone at the bench shape: CUDA-event time back to back (working set cycling through L2) and with
the L2 flushed before every launch.  Prints the algorithmic GB/s (68.2 MB per launch)."""
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
eng.decode_greedy(ws, d, 2)
torch.cuda.synchronize()
ALG = 4 * 512 * (2 * 26 * 8 + 4 * 26) * 64
for _ in range(5):
    eng.attention(ws, d)
torch.cuda.synchronize()
n = 50
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    eng.attention(ws, d)
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) / n * 1e3
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
tot = 0.0
for _ in range(10):
    flush.zero_()
    a.record()
    eng.attention(ws, d)
    b.record()
    torch.cuda.synchronize()
    tot += a.elapsed_time(b)
env = ' '.join('%s=%s' % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith('STAT_'))
print('attention: back-to-back %.2f us (%.0f GB/s)   flushed %.2f us (%.0f GB/s)  [%s]' % (
    us, ALG / us / 1e3, tot / 10 * 1e3, ALG / (tot / 10 * 1e3) / 1e3, env))
