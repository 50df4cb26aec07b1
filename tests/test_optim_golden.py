"""The optimizer oracle is pinned to the reference's own adam / adadelta source (common.py:178-230): the
trajectories in tests/golden/optim_ref.npz were produced by executing that source text
(tests/golden/make_optim_golden.py); here the restatement oracle/optim_oracle.py must reproduce them, and --
where /root/reference is mounted -- the fixture is regenerated live and compared."""
import os

import numpy as np
import pytest

from oracle import optim_oracle as oo, ref_exec

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'optim_ref.npz')


def _replay(z, name):
    n = z['p0'].shape[0]
    opt = oo.Adam(n) if name == 'adam' else oo.Adadelta(n)
    p = z['p0'].copy()
    out = []
    for g in z['grads']:
        if name == 'adadelta':
            opt.grad_shared(g)
        p = opt.update(p, g)
        out.append(p.copy())
    return np.stack(out)


@pytest.mark.parametrize('name', ['adam', 'adadelta'])
def test_oracle_reproduces_reference_optimizer(name):
    z = np.load(GOLD)
    got = _replay(z, name)
    # same float32 operations in the same order; numpy may fuse nothing, so agreement is to the last bits
    np.testing.assert_allclose(got, z[name], rtol=2e-6, atol=1e-9)
    assert np.abs(z[name][-1] - z['p0']).max() > 1e-4          # the parameters did move


@pytest.mark.skipif(not ref_exec.available(), reason='reference sources not mounted')
def test_fixture_matches_live_reference_source():
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_optim_golden', os.path.join(os.path.dirname(GOLD),
                                                                                   'make_optim_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    live = mod.trajectories()
    z = np.load(GOLD)
    for k in ('p0', 'grads', 'adam', 'adadelta'):
        assert np.array_equal(live[k], z[k]), k
