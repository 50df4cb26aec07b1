"""Graph-replay time of the beam search at the bench's secondary workload (32 clips, k=5, maxlen 30, K0 included)."""
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
B, k, L = int(os.environ.get('BEAM_B', 32)), int(os.environ.get('BEAM_K', 5)), 30
devf = [torch.from_numpy(a[:B]).cuda() for a in feats]
for _ in range(3):
    eng.beam_captions(*devf, k=k, maxlen=L, use_graph=True)
torch.cuda.synchronize()
n = 10
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    out = eng.beam_captions(*devf, k=k, maxlen=L, use_graph=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
env = ' '.join('%s=%s' % (kk, os.environ[kk]) for kk in sorted(os.environ) if kk.startswith('STAT_'))
print('beam B=%d k=%d maxlen=%d: %.3f ms per batch, %.0f captions/s  [%s]' % (B, k, L, ms, B / ms * 1e3, env))
