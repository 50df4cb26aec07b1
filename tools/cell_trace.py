"""globaltimer stamps (ns) of thread 0 of every CTA of the last cell_kernel launch of a short greedy decode."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
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
eng.decode_greedy(ws, d, 4)
torch.cuda.synchronize()
buf = torch.zeros(16384, dtype=torch.int64, device='cuda')
os.environ['STAT_ATT_TRACE_OFF'] = '1'
lib.stat_debug_gemm_trace(C.c_void_p(buf.data_ptr()))
torch.cuda._sleep(20000000)
eng.decode_greedy(ws, d, 4)
torch.cuda.synchronize()
lib.stat_debug_gemm_trace(C.c_void_p(0))
t = buf.cpu().numpy()[4096:4096 + 148 * 16].reshape(148, 16).astype(np.float64)
t0 = t[:, 0][t[:, 0] > 0].min()
names = ['start', 'prologue', 'pdl_wait', 'W1 here', 'X1 here', 'compute1', 'epilogue1', 'barrier', 'W2 here', 'X2 here',
         'compute2', 'end']
for i, n in enumerate(names):
    col = t[:, i]
    col = col[col > 0] - t0
    if len(col):
        print('%-10s min %8.0f  median %8.0f  max %8.0f ns   (%d CTAs)' % (n, col.min(), np.median(col), col.max(), len(col)))
