"""Why is the end-to-end H2D rate 50.8 GB/s when one raw copy runs at 55?  (a) the copy loop of caption_stream alone,
(b) caption_stream itself, (c) the copy loop with graph replays on resident inputs beside it (no staging copy)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import synthetic
from video_description_with_spatial_temporal_attention_b200.engine import Engine
import bench

bench.bind_to_gpu_numa_node(0)
o = stat.baseline_options()
params = synthetic.trained_like_params(o, seed=7)
feats = bench.make_inputs(o, 1234)
eng = Engine(o)
eng.set_params(params)
host = [torch.from_numpy(a).pin_memory() for a in feats]
devf = [h.cuda() for h in host]
nbytes = sum(h.numel() * 4 for h in host)
eng.greedy_captions(*devf, maxlen=20, use_graph=True)
st = eng.graph_inputs(bench.B, bench.T, bench.R, 20)
static = [st['ctxg'], st['mask'], st['ctxl'], st['ctxm']]
stage = [[torch.empty_like(s) for s in static] for _ in range(2)]
copy = torch.cuda.Stream()
N = 20


def timed(fn):
    fn(3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn(N)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / N


def copies_only(n):
    cur = torch.cuda.current_stream()
    copy.wait_stream(cur)
    with torch.cuda.stream(copy):
        for i in range(n):
            for d, s in zip(stage[i % 2], host):
                d.copy_(s, non_blocking=True)
    cur.wait_stream(copy)


def copies_beside_replays(n):
    cur = torch.cuda.current_stream()
    copy.wait_stream(cur)
    with torch.cuda.stream(copy):
        for i in range(n):
            for d, s in zip(stage[i % 2], host):
                d.copy_(s, non_blocking=True)
    for i in range(3 * n):
        eng.greedy_captions(*static, maxlen=20)
    cur.wait_stream(copy)


def stream(n):
    for _ in eng.caption_stream((host for _ in range(n)), 20):
        pass


for name, fn in (('copy loop alone', copies_only), ('copy loop beside graph replays (no staging copy)', copies_beside_replays),
                 ('caption_stream', stream)):
    ms = timed(fn)
    print('%-50s %.3f ms per batch  %.1f GB/s' % (name, ms, nbytes / ms / 1e6))
