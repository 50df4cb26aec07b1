"""CPU-only checks (no GPU in the authoring container): the C-ABI library loads and
exports every symbol include/stat_b200.h declares, the host mirror's parameter
initialiser is bit-identical with the reference's, and the host-side beam search
reproduces the reference's hypotheses when driven by the oracle's f_init/f_next."""
import ctypes
import os
import re

import numpy as np
import pytest

import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import _lib, common, model_attention as ma
from oracle import stat_oracle as so
from tests.golden_util import NAMES, Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, 'include', 'stat_b200.h')) as fh:
        src = fh.read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(stat_[a-z_0-9]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_lib.LIB_PATH):
        from video_description_with_spatial_temporal_attention_b200 import build
        build.build_lib()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), s
    assert _lib.load().stat_version() == 1


def test_sizes_and_argument_errors_without_gpu():
    lib = _lib.load()
    o = stat.baseline_options()
    d = _lib.StatDims(B=64, T=26, R=8, Dg=o['ctxg_dim'], Dm=o['ctxm_dim'], Dr=o['ctxl_dim'], H=o['dim'],
                      E=o['dim_word'], V=o['n_words'], flags=_lib.flags_of(o))
    nprep = lib.stat_prepared_bytes(ctypes.byref(d))
    nws = lib.stat_workspace_bytes(ctypes.byref(d), 64)
    # packed weights: the 25 M parameters + the (V+1,4H) token table + a Wemb copy
    assert 200e6 < nprep < 300e6
    # 7 context blocks of the batch dominate: 64 * (3*8 + 4) * 26 * 512 * 4 B = 95 MB
    assert 95e6 < nws < 130e6
    off, nb = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.stat_workspace_region(ctypes.byref(d), 64, b'pctxl', ctypes.byref(off), ctypes.byref(nb)) == 0
    assert nb.value == 64 * 26 * 8 * 512 * 4
    assert lib.stat_workspace_region(ctypes.byref(d), 64, b'nope', ctypes.byref(off), ctypes.byref(nb)) != 0
    assert b'unknown region' in lib.stat_last_error()
    # the reference graph's Dg == H constraint is enforced at the boundary
    bad = _lib.StatDims(B=1, T=2, R=2, Dg=16, Dm=8, Dr=8, H=8, E=8, V=11, flags=0)
    assert lib.stat_prepared_bytes(ctypes.byref(bad)) == 0
    assert b'ctxg_dim == dim' in lib.stat_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from video_description_with_spatial_temporal_attention_b200.engine import Engine
    with pytest.raises(_lib.StatError):
        Engine(stat.default_options())


def test_init_params_bit_identical_with_reference_fixture():
    g = Golden('ref_tiny_init')
    common.rng_numpy.seed(1234)
    p = ma.Attention().init_params(g.options)
    assert list(p) == list(g.params)
    for k in p:
        assert np.asarray(p[k]).dtype == np.float32
        np.testing.assert_array_equal(np.asarray(p[k]), g.params[k])
    o2 = stat.default_options(dim=8, dim_word=8, ctxg_dim=20, global_proj=True, n_words=11, ctxl_dim=12,
                              ctxm_dim=10)
    p2 = ma.Attention().init_params(o2)
    assert len(p2) == 43 and list(p2)[5:7] == ['ff_global_W', 'ff_global_b']
    with pytest.raises(ValueError):
        ma.Attention().init_params(dict(o2, global_proj=False))


@pytest.mark.parametrize('name', NAMES)
def test_host_beam_search_on_oracle_callables(name):
    """gen_sample's host bookkeeping (shrinking beam, retirement on eos, maxlen
    survivors) against the hypotheses the reference produced."""
    g = Golden(name)
    x, m, ctxg, mg, ctxl, ml, ctxm, mm = g.batch
    f_init, f_next = so.make_sampler(g.params, g.options, hoist=True)
    model = ma.Attention()
    for k in g.ks:
        for b in range(ctxg.shape[0]):
            want, want_sc = g.hyps(k, b)
            got, got_sc, _, _ = model.gen_sample(None, f_init, f_next, ctxg[b], mg[b], ctxl[b], ml[b], ctxm[b],
                                                 mm[b], g.options, None, k, g.maxlen, False)
            assert got == want, (name, k, b)
            np.testing.assert_allclose(np.asarray(got_sc), want_sc, atol=2e-5)


def test_shared_and_function_plumbing():
    o = stat.default_options(dim=8, dim_word=8, ctxg_dim=8, n_words=11, ctxl_dim=12, ctxm_dim=12)
    model = ma.Attention()
    tp = model.init_tparams(model.init_params(o))
    v0 = tp.version
    tp['Wemb'].set_value(tp['Wemb'].get_value() * 2)
    assert tp.version == v0 + 1
    r = model.build_model(tp, o)
    assert len(r) == 16
    use_noise, cost = r[1], r[14]
    use_noise.set_value(1.)
    assert float(use_noise.get_value()) == 1.0
    f = ma.function(list(r[2:10]), -cost)
    assert callable(f)
    with pytest.raises(TypeError):
        f(1, 2, 3)
    assert common.generate_minibatch_idx(10, 4) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
