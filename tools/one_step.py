"""One eager pass of the bench workload (K0 + 20 greedy steps at B=64) for ncu."""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(n):
    ws, d = eng.precompute(*devf)
    toks, lens, scores = eng.decode_greedy(ws, d, bench.MAXLEN)
torch.cuda.synchronize()
print('ok', int(toks.sum()))
