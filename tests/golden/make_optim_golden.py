"""Fixture for the optimizers: runs the reference's OWN adam / adadelta source text (common.py:178-230,
executed on oracle/mini_theano via oracle/ref_exec._load) for a few updates on seeded gradients and stores
the parameter trajectories.  Authoring container only (needs /root/reference):

    python tests/golden/make_optim_golden.py        # -> tests/golden/optim_ref.npz
"""
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_exec  # noqa: E402


def trajectories(n=257, steps=6, seed=3):
    th, common, _ = ref_exec._load()
    tt = th.tensor
    rng = np.random.RandomState(seed)
    p0 = rng.randn(n).astype('float32')
    grads = [(rng.randn(n) * (5.0 if s % 2 else 0.02)).astype('float32') for s in range(steps)]
    out = {'p0': p0, 'grads': np.stack(grads)}
    for name in ('adam', 'adadelta'):
        tparams = OrderedDict(p=th.shared(p0.copy(), name='p'))
        g_in = tt.vector('g', dtype='float32')
        lr = tt.vector('lr', dtype='float32')          # unused by the update rules (0-d scalars have no ctor here)
        f_grad_shared, f_update = getattr(common, name)(lr, tparams, [g_in], [g_in], g_in.sum(), [])
        traj = []
        for g in grads:
            f_grad_shared(g)
            f_update(np.zeros(1, 'float32'))
            traj.append(np.asarray(tparams['p'].get_value(), 'float32').copy())
        out[name] = np.stack(traj)
    return out


if __name__ == '__main__':
    o = trajectories()
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'optim_ref.npz'), **o)
    print({k: v.shape for k, v in o.items()})
