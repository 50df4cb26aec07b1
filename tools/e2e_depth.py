"""End-to-end captions/s of Engine.caption_stream for several pipeline depths."""
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
host = [torch.from_numpy(a).pin_memory() for a in feats]


def batches(n):
    for _ in range(n):
        yield host


for depth in (2, 3, 4, 2):
    for _ in eng.caption_stream(batches(3), bench.MAXLEN, depth=depth):
        pass
    torch.cuda.synchronize()
    n = 30
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in eng.caption_stream(batches(n), bench.MAXLEN, depth=depth):
        pass
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print('depth %d: %.3f ms per batch -> %.0f captions/s (%.1f GB/s of H2D)' % (depth, ms, 64 / ms * 1e3, 259004928 / ms / 1e6))
