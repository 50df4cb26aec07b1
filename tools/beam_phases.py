"""Per-phase device time of the beam search (B=32 clips, k=5, maxlen=30) from the library's event profile."""
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
B, k, L = int(os.environ.get('BEAM_B', 32)), 5, 30
devf = [torch.from_numpy(a[:B]).cuda() for a in feats]
eng.beam_captions(*devf, k=k, maxlen=L)
torch.cuda.synchronize()


def run():
    torch.cuda._sleep(20000000)
    eng.beam_captions(*devf, k=k, maxlen=L)


ph = eng.profile(run)
tot = 0.0
for name, (ms, cnt) in ph.items():
    if cnt:
        print('%-16s %8.3f ms  %4d groups  %7.1f us each' % (name, ms, cnt, ms / cnt * 1e3))
        tot += ms
print('sum %.3f ms' % tot)
