"""Per-(CTA, group) clock stamps of att_group_kernel (thread 0 of every group):
[globaltimer, clk start, clk constants loaded, (P arrived, V arrived, frame done) x <=4, ..., clk end]."""
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
devf = [torch.from_numpy(a).cuda() for a in feats]
ws, d = eng.precompute(*devf)
eng.decode_greedy(ws, d, 1)
torch.cuda.synchronize()
lib = eng.lib
lib.stat_debug_gemm_trace.argtypes = [C.c_void_p]
NCTA, G = 128, 4
buf = torch.zeros(NCTA * G * 16 + 1024, dtype=torch.int64, device='cuda')
flush = torch.empty(96 * 1024 * 1024, dtype=torch.float32, device='cuda')
for it in range(4):
    buf.zero_()
    if it >= 2:
        flush.zero_()
    lib.stat_debug_gemm_trace(C.c_void_p(buf.data_ptr()))
    eng.attention(ws, d)
    torch.cuda.synchronize()
    lib.stat_debug_gemm_trace(C.c_void_p(0))
    t = buf.cpu().numpy()[:NCTA * G * 16].reshape(NCTA, G, 16)
    g0 = t[:, :, 0].min()
    print('run', it, '(cold L2)' if it >= 2 else '(warm L2)')
    print('  CTA start skew (globaltimer ns): min 0 max %d' % (t[:, :, 0].max() - g0))
    rel = t[:, :, 1:] - t[:, :1, 1:2]
    names = ['start', 'const', 'P0', 'V0', 'F0', 'P1', 'V1', 'F1', 'P2', 'V2', 'F2', 'P3', 'V3', 'F3', 'end']
    for c in (0, 1, 64, 127):
        for g in range(G):
            print('  cta %3d g%d ' % (c, g) + ' '.join('%s=%d' % (n, v) for n, v in zip(names, rel[c, g]) if v > -10**9 and (v != -t[c, 0, 1])))
    m = np.where(t[:, :, 1:] != 0, rel, -1)
    print('  median over CTAs (cycles): ' + ' '.join('%s=%d' % (n, np.median(m[:, 0, k])) for k, n in enumerate(names)))
    print('  max    over CTAs (cycles): ' + ' '.join('%s=%d' % (n, m[:, :, k].max()) for k, n in enumerate(names)))
