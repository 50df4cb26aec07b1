"""Batch time of the greedy decode: eager launches (host enqueued ahead of a parked GPU) vs graph replay."""
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
L = 20
for mode in ('eager', 'graph'):
    for _ in range(3):
        eng.greedy_captions(*devf, maxlen=L, use_graph=(mode == 'graph'))
    torch.cuda.synchronize()
    n = 5
    torch.cuda._sleep(40000000)          # ~20 ms: the host gets ahead of the GPU
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        eng.greedy_captions(*devf, maxlen=L, use_graph=(mode == 'graph'))
    b.record()
    torch.cuda.synchronize()
    print('%s: %.3f ms per batch (STAT_PDL=%s STAT_OVERLAP=%s)' % (mode, a.elapsed_time(b) / n,
                                                                   os.environ.get('STAT_PDL', '1'),
                                                                   os.environ.get('STAT_OVERLAP', '1')))
