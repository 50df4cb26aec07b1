"""Fixture for the training step: cost and the 41(+2) gradients of the toy cases of tests/test_backward_emu.py /
tests/test_gpu_train.py from the fp64 gradient oracle (oracle/grad_oracle.py: torch autograd over the restated
forward, which is itself pinned to the reference's source through tests/golden/ref_*.npz).  Freezes the oracle's
output so that a later edit of the oracle cannot silently move the target the CUDA backward pass is compared with.

    python tests/golden/make_grad_golden.py        # -> tests/golden/grad_toy.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import grad_oracle as go  # noqa: E402
from tests.test_backward_emu import _case  # noqa: E402

KW = dict(alpha_c=0.70602, decay_c=1e-4)


def build():
    out = {}
    for gp in (False, True):
        o, params, batch = _case(gp)
        cost, grads, ex = go.cost_and_grads(params, o, batch, **KW)
        tag = 'gp%d' % int(gp)
        out[tag + '/cost'] = np.float64(cost)
        out[tag + '/g2'] = np.float64(ex['g2'])
        for k, v in grads.items():
            out[tag + '/' + k] = np.asarray(v, 'float64')
    return out


if __name__ == '__main__':
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'grad_toy.npz')
    np.savez_compressed(path, **build())
    print(path, os.path.getsize(path), 'bytes')
