"""Phase stamps (clock64) of one CTA of the tensor-core GEMM for a few skinny shapes."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200.engine import Engine

eng = Engine(stat.default_options())
lib = eng.lib
lib.stat_debug_gemm_trace.argtypes = [C.c_void_p]
buf = torch.zeros(256, dtype=torch.int64, device='cuda')
for (M, N, K, SW) in [(64, 12594, 512, True), (13312, 512, 4096, False)]:
    A = torch.randn(M, K, device='cuda')
    Bt = torch.randn(N, K, device='cuda')
    for it in range(3):
        buf.zero_()
        lib.stat_debug_gemm_trace(C.c_void_p(buf.data_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.gemm(A, Bt, swap=SW)
        e1.record()
        torch.cuda.synchronize()
        lib.stat_debug_gemm_trace(C.c_void_p(0))
    t = buf.cpu().tolist()
    t0 = t[0]
    rel = lambda i: (t[i] - t0) if t[i] else None
    print('shape', (M, N, K), 'event_us %.1f' % (e0.elapsed_time(e1) * 1e3))
    print('  setup', rel(1), 'full', [rel(2 + k) for k in range(16)])
    print('  split_done', [rel(20 + k) for k in range(16)])
    print('  mma_start', [rel(40 + k) for k in range(16)])
    print('  tmem_full', rel(60), 'epi_done', rel(61), 'end', rel(62))
    print('  epi', [rel(64 + k) for k in range(16)])
