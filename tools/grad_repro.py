"""Bit-reproducibility hunt: the config-3 gradient (B=128) N times in one process; which tensors ever differ from run 0."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import synthetic
from video_description_with_spatial_temporal_attention_b200.train import Trainer
from oracle import stat_oracle as so

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
noise = len(sys.argv) > 2 and sys.argv[2] == 'noise'
o = stat.baseline_options()
params = so.trained_like_params(o, seed=12)
batch = synthetic.make_batch(o, B=128, T=26, R=8, L=20, seed=12, zero_tail=True)
tr = Trainer(params, o, use_noise=noise, alpha_c=0.70602, decay_c=0.0)
ref = None
bad = {}
costs = []
for i in range(N):
    if noise:
        tr.gen.manual_seed(1234) if hasattr(tr, 'gen') else None
    c = tr.f_grad_shared(*batch)[0]
    costs.append(float(c))
    g = {k: v.copy() for k, v in tr.grads().items()}
    if ref is None:
        ref = g
        continue
    for k in g:
        if not np.array_equal(g[k], ref[k]):
            d = np.abs(g[k].astype('float64') - ref[k])
            bad.setdefault(k, []).append((i, float(d.max()), int((d > 0).sum()), bool(np.isfinite(g[k]).all())))
print('runs', N, 'distinct costs', len(set(costs)), sorted(set(costs))[:4])
for k, v in bad.items():
    print('%-22s differs in %d runs, e.g. run %d: max|d| %.3g over %d elements, finite %s' % (k, len(v), v[0][0], v[0][1], v[0][2], v[0][3]))
print('non-reproducible tensors:', len(bad))
