"""The gradient oracle for the (not yet built) CUDA backward pass: its forward must equal the numpy
oracle that is pinned to the reference, its gradients central differences of its own cost."""
import numpy as np

import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import synthetic
from oracle import grad_oracle as go, stat_oracle as so


def _case(global_proj):
    kw = dict(dim=12, dim_word=10, ctxl_dim=14, ctxm_dim=9, n_words=17)
    o = stat.default_options(ctxg_dim=20, global_proj=True, **kw) if global_proj else \
        stat.default_options(ctxg_dim=12, **kw)
    params = so.trained_like_params(o, seed=11)
    batch = synthetic.make_batch(o, B=4, T=5, R=3, L=6, seed=11, zero_tail=True)
    return o, params, batch


def test_forward_equals_numpy_oracle():
    for gp in (False, True):
        o, params, batch = _case(gp)
        want, allv = so.forward_teacher(params, o, *batch, dtype=np.float64, return_all=True)
        cost, grads, ex = go.cost_and_grads(params, o, batch)
        np.testing.assert_allclose(ex['logp'], want, rtol=0, atol=1e-11)
        assert abs(cost - (-want).mean()) < 1e-11
        for k, kk in (('alphals', 'alphaL'), ('alphags', 'alphaG'), ('alphams', 'alphaM'), ('alphalts', 'alphaLT')):
            np.testing.assert_allclose(ex[k], np.stack([s[kk] for s in allv['steps']]), atol=1e-12)
        assert list(grads.keys()) == list(params.keys())                 # the reference's gradient order
        assert all(grads[k].shape == np.asarray(params[k]).shape for k in params)


def test_gradients_match_central_differences():
    o, params, batch = _case(True)
    kw = dict(alpha_c=0.70602, decay_c=1e-4)
    cost, grads, _ = go.cost_and_grads(params, o, batch, **kw)
    rng = np.random.RandomState(3)
    eps = 1e-6
    for k in ('Wemb', 'ff_local_W', 'ff_state_b', 'decoder_U', 'decoder_Wc', 'decoder_Wclt_att', 'decoder_Wdl_att',
              'decoder_Ul_att', 'decoder_clt_att', 'decoder_W_sel', 'decoder_b_sel', 'ff_logit_ctxglm_W', 'ff_logit_b',
              'ff_global_W', 'decoder_bl_att'):
        v = np.asarray(params[k], 'float64')
        direction = rng.randn(*v.shape) if v.shape else np.float64(1.0)
        plus, minus = dict(params), dict(params)
        plus[k] = v + eps * direction
        minus[k] = v - eps * direction
        cp = go.cost_and_grads(plus, o, batch, **kw)[0]
        cm = go.cost_and_grads(minus, o, batch, **kw)[0]
        fd = (cp - cm) / (2 * eps)
        an = float((grads[k] * direction).sum())
        assert abs(fd - an) <= 2e-6 * max(1.0, abs(an)), (k, fd, an)


def test_clip_and_regularisers():
    from oracle import optim_oracle as oo
    o, params, batch = _case(False)
    c0, g0, ex0 = go.cost_and_grads(params, o, batch)
    c1, g1, ex1 = go.cost_and_grads(params, o, batch, alpha_c=0.5, decay_c=1e-3)
    wd = 1e-3 * sum(float((np.asarray(v, 'float64') ** 2).sum()) for v in params.values())
    reg = sum(0.5 * float(((1. - ex0[k].sum(0)) ** 2).sum(0).mean()) for k in ('alphags', 'alphals', 'alphams', 'alphalts'))
    assert abs(c1 - (c0 + wd + reg)) < 1e-10
    # global-norm clipping as model_attention.py:1194-1203 == the flat-buffer rule the optimizer kernels use
    flat = np.concatenate([g.reshape(-1) for g in g0.values()])
    clip_c = 0.5 * float(np.sqrt((flat ** 2).sum()))
    _, gc, exc = go.cost_and_grads(params, o, batch, clip_c=clip_c)
    want, g2 = oo.clip(flat.astype('float32'), clip_c)
    got = np.concatenate([g.reshape(-1) for g in gc.values()])
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-9)
    assert abs(exc['g2'] - float(g2)) < 1e-5 * exc['g2']


def test_oracle_reproduces_the_committed_gradient_fixture():
    """tests/golden/grad_toy.npz (tests/golden/make_grad_golden.py) freezes the target of the CUDA backward pass."""
    import os
    from tests.golden import make_grad_golden as mg
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'grad_toy.npz')
    want = np.load(path)
    got = mg.build()
    assert sorted(want.files) == sorted(got.keys())
    for k in want.files:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-9, atol=1e-12, err_msg=k)
