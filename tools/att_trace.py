"""clock64 stamps of CTA 0 / thread 0 of att_stream_kernel: per frame
[loop top, full-wait done, after barrier 1 (A), after barrier 2 (C), frame done]."""
import ctypes as C
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
lib = eng.lib
lib.stat_debug_gemm_trace.argtypes = [C.c_void_p]
buf = torch.zeros(128, dtype=torch.int64, device='cuda')
for it in range(3):
    buf.zero_()
    lib.stat_debug_gemm_trace(C.c_void_p(buf.data_ptr()))
    eng.attention(ws, d)
    torch.cuda.synchronize()
    lib.stat_debug_gemm_trace(C.c_void_p(0))
    t = buf.cpu().tolist()
    t0 = t[0]
    print('run', it, 'kernel stamps: start 0, state loaded %d, loop end %d, cluster sync %d, merged %d, exit %d' % tuple(t[64 + k] - t[64] for k in range(1, 6)), 'first frame top', t[0] - t[64])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.attention(ws, d)
    e1.record()
    torch.cuda.synchronize()
    print('  10 back-to-back launches: %.1f us each' % (e0.elapsed_time(e1) * 100))
