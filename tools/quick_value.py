"""Graph-replay time of the bench workload (B=64 greedy decode, K0 included) for two caption lengths: the
difference gives the device time of one decode step.  One line per run; environment switches are echoed.

  VARS="STAT_ATT_SERP STAT_SIDE_PRIO" python tools/quick_value.py
"""
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
res = {}
for L in (2, 20):
    eng.greedy_captions(*devf, maxlen=L, use_graph=True)
    st = eng.graph_inputs(bench.B, bench.T, bench.R, L)
    static = [st['ctxg'], st['mask'], st['ctxl'], st['ctxm']]
    for _ in range(5):
        eng.greedy_captions(*static, maxlen=L)
    torch.cuda.synchronize()
    n = 30
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        out = eng.greedy_captions(*static, maxlen=L)
    b.record()
    torch.cuda.synchronize()
    res[L] = a.elapsed_time(b) / n
    if L == 20:
        chk = int(out[0].clamp(min=0).sum().item())
ws, d = eng.precompute(*devf)
def eager():
    eng.decode_greedy(ws, d, 20)
eager()
torch.cuda.synchronize()
def parked():
    torch.cuda._sleep(12000000)
    eager()
acc = {}
for _ in range(3):
    for k, (m, c) in eng.profile(parked).items():
        acc[k] = acc.get(k, 0.0) + m / 3
phases = ' '.join('%s=%.1f' % (k.replace('step_', ''), v / 20 * 1e3) for k, v in acc.items() if k.startswith('step_'))
env = ' '.join('%s=%s' % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith('STAT_'))
print('L20 %.4f ms  L2 %.4f ms  per-step %.2f us  captions/s %.0f  tokens_checksum %d  [%s]\n    eager us/step: %s' % (
    res[20], res[2], (res[20] - res[2]) / 18 * 1e3, bench.B / res[20] * 1e3, chk, env, phases))
