"""Synthetic inputs of the MSVD shape.

The reference's features are not shipped (data_engine.py:40-56 reads HDF5 files
from the authors' disks), so every test and benchmark draws inputs here, laid
out exactly like ``data_engine.prepare_data``'s 8-tuple (data_engine.py:258-337):
``x (L,B) int64, x_mask (L,B) f32, ctxg (B,T,Dg), ctxg_mask (B,T),
ctxl (B,T,R,Dr), ctxl_mask (B,T,R), ctxm (B,T,Dm), ctxm_mask (B,T)``.
"""
from __future__ import annotations

import numpy as np


def feature_masks(ctxg, ctxl, ctxm):
    """Masks as data_engine.get_ctx*_mask computes them (data_engine.py:169-219):
    a slot is valid when its feature vector does not sum to zero."""
    f = lambda a: (a.sum(axis=-1) != 0).astype('int32').astype('float32')
    return f(ctxg), f(ctxl), f(ctxm)


def make_features(B, T, R, Dg, Dr, Dm, seed=1234, nonneg=False, zero_tail=False):
    r = np.random.RandomState(seed)
    def draw(*s):
        a = r.randn(*s).astype('float32')
        return np.abs(a) if nonneg else 0.5 * a
    ctxg, ctxl, ctxm = draw(B, T, Dg), draw(B, T, R, Dr), draw(B, T, Dm)
    if zero_tail:
        # zero-padded trailing frames (data_engine.py:83-91) exercise `counts`
        for b in range(B):
            n = int(r.randint(0, min(7, T)))
            if n:
                ctxg[b, T - n:] = 0
                ctxl[b, T - n:] = 0
                ctxm[b, T - n:] = 0
    mg, ml, mm = feature_masks(ctxg, ctxl, ctxm)
    return ctxg, mg, ctxl, ml, ctxm, mm


def make_captions(B, V, L, seed=1234, ragged=True):
    """x[:len_b] = tokens in [2,V), x[len_b:] = 0 (eos + padding);
    mask[:len_b+1] = 1 (data_engine.py:331-335)."""
    r = np.random.RandomState(seed + 1)
    if ragged:
        lens = np.clip(r.poisson(7, size=B) + 1, 1, L - 1)
        lens[r.randint(B)] = L - 1
    else:
        lens = np.full((B,), L - 1)
    x = np.zeros((L, B), 'int64')
    m = np.zeros((L, B), 'float32')
    for b in range(B):
        x[:lens[b], b] = r.randint(2, V, size=lens[b])
        m[:lens[b] + 1, b] = 1.
    return x, m


def make_batch(options, B, T, R, L, seed=1234, ragged=True, nonneg=False,
               zero_tail=False):
    o = options
    ctxg, mg, ctxl, ml, ctxm, mm = make_features(
        B, T, R, o['ctxg_dim'], o['ctxl_dim'], o['ctxm_dim'], seed, nonneg, zero_tail)
    x, m = make_captions(B, o['n_words'], L, seed, ragged)
    return x, m, ctxg, mg, ctxl, ml, ctxm, mm


def trained_like_params(options, seed=7):
    """The reference's parameter dict (init_params key order and shapes) filled with
    magnitudes resembling a trained model, so attention and vocabulary softmaxes are
    far from uniform.  Random-init stand-in for the checkpoints that are not shipped
    (bench / smoke input only)."""
    from collections import OrderedDict
    from . import common
    from .model_attention import Attention
    state = common.rng_numpy.get_state()
    common.rng_numpy.seed(1234)
    try:
        p = Attention().init_params(options)
    finally:
        common.rng_numpy.set_state(state)
    r = np.random.RandomState(seed)
    H = options['dim']
    out = OrderedDict()
    for k, v in p.items():
        v = np.asarray(v, 'float32')
        if v.ndim == 2 and v.shape[1] == 1:
            out[k] = (r.randn(*v.shape) * (2.0 / np.sqrt(H))).astype('float32')
        elif v.ndim == 2:
            if k == 'Wemb':
                out[k] = (r.randn(*v.shape) * 0.5).astype('float32')
                continue
            gain = 3.0 if k == 'ff_logit_W' else (2.0 if k.split('_')[1] in ('local', 'motion', 'global', 'state',
                                                                             'memory') and k.startswith('ff_') else 1.0)
            out[k] = (r.randn(*v.shape) * (gain / np.sqrt(v.shape[0]))).astype('float32')
        elif v.ndim == 1:
            out[k] = (r.randn(*v.shape) * 0.1).astype('float32')
        else:
            out[k] = np.float32(r.randn() * 0.1)
    return out
