"""Clock stamps of CTA 0 of the fused step kernels (B: gates, C: queries + readout, logits) at the bench shape.
    python tools/fused_trace.py            -> per phase: k-atom arrival / split / MMA stamps in cycles"""
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
lib = eng.lib
lib.stat_debug_gemm_trace.argtypes = [C.c_void_p]
devf = [torch.from_numpy(a).cuda() for a in feats]
ws, d = eng.precompute(*devf)
eng.decode_greedy(ws, d, 3)
torch.cuda.synchronize()
buf = torch.zeros(256, dtype=torch.int64, device='cuda')
# the trace buffer is overwritten by every fused launch: run a 1-step decode and keep the stamps of the LAST launch
# of each kind by running with maxlen=1 (launch order: C(queries) , att, B, C(z), logits)
for kind, label in ((1, 'B: gates tiles (swap, bq=32, 32 k-atoms)'), (0, 'C: queries tiles (swap, bq=32, 16 k-atoms)'),
                    (4, 'logits (normal tiles, bq=128, 16 k-atoms)')):
    maxlen = 2
    os.environ['STAT_TRACE_KIND'] = str(kind)
    buf.zero_()
    lib.stat_debug_gemm_trace(C.c_void_p(buf.data_ptr()))
    eng.decode_greedy(ws, d, maxlen)
    torch.cuda.synchronize()
    lib.stat_debug_gemm_trace(C.c_void_p(0))
    t = buf.cpu().tolist()
    t0 = t[0]
    rel = lambda i: (t[i] - t0) if t[i] else None
    print(label)
    print('  setup', rel(1), 'acc_ready', rel(140), 'epi_done', rel(141), 'end', rel(142))
    print('  epilogue: stage start', rel(148), 'staged', rel(149), 'done', rel(154))
    print('  tma_issue ', [rel(100 + k) for k in range(32)])
    print('  full      ', [rel(2 + k) for k in range(32)])
    print('  split_done', [rel(36 + k) for k in range(32)])
    print('  mma_issue ', [rel(180 + k) for k in range(32)])
