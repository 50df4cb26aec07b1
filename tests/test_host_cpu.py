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
    # a beam search over 160 rows (32 clips x k = 5) keeps the k-split planes of the skinny products: its workspace
    # holds 160 rows of them (the products are issued as a 128-row and a 32-row launch), 200 rows fall back to one plane
    w128 = lib.stat_workspace_bytes(ctypes.byref(d), 128)
    w160 = lib.stat_workspace_bytes(ctypes.byref(d), 160)
    w200 = lib.stat_workspace_bytes(ctypes.byref(d), 200)
    assert w128 < w160 and (w160 - w128) / 32 > (w200 - w160) / 40
    # process-wide switches validate their argument (no GPU needed)
    assert lib.stat_set_beam_share(2) != 0 and b'beam share' in lib.stat_last_error()
    assert lib.stat_set_beam_share(0) == 0 and lib.stat_set_beam_share(-1) == 0
    assert lib.stat_set_step_impl(3) != 0
    assert lib.stat_set_step_impl(-1) == 0


def test_training_entry_points_without_gpu():
    """Sizes, phase names and argument errors of the backward-pass ABI (no compute without a GPU)."""
    lib = _lib.load()
    o = stat.baseline_options()
    d = _lib.StatDims(B=128, T=26, R=8, Dg=o['ctxg_dim'], Dm=o['ctxm_dim'], Dr=o['ctxl_dim'], H=o['dim'],
                      E=o['dim_word'], V=o['n_words'], flags=_lib.flags_of(o))
    n20 = lib.stat_grad_workspace_bytes(ctypes.byref(d), 20)
    n10 = lib.stat_grad_workspace_bytes(ctypes.byref(d), 10)
    # transposed raw local features (4096 x 26624 floats) and their H-wide partner dominate; a few GB at most
    assert 1.5e9 < n10 < n20 < 4e9
    assert lib.stat_grad_workspace_bytes(ctypes.byref(d), 0) == 0
    names = [lib.stat_grad_profile_phase_name(i).decode() for i in range(lib.stat_grad_profile_phases())]
    assert 'loop_att_main' in names and 'readout_backward' in names and len(set(names)) == len(names)
    assert lib.stat_grad_profile_phase_name(len(names)) is None
    sp, blocks = _lib.StatParams(), _lib.StatFwdBlocks()
    rc = lib.stat_grad_shared(ctypes.byref(d), ctypes.byref(sp), ctypes.byref(blocks), 20, *([None] * 14), 1.0, 0.0, 0.0,
                              ctypes.byref(sp), None, None)
    assert rc == -1 and b'NULL' in lib.stat_last_error()
    with pytest.raises(_lib.StatError):
        _lib.check(rc)


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


# ---------------------------------------------------------------------------
# minibatch layout (data_engine.py:258-337) and checkpoint files (N3, N4)
# ---------------------------------------------------------------------------
def test_prepare_data_layout_matches_reference_rules():
    import video_description_with_spatial_temporal_attention_b200 as stat
    from video_description_with_spatial_temporal_attention_b200 import data_engine as de
    o = stat.default_options(dim=16, dim_word=16, ctxg_dim=16, ctxl_dim=24, ctxm_dim=20, n_words=30)
    eng = de.synthetic_engine(o, n_videos=6, caps_per_video=2, T=5, R=3, seed=3)
    ids = eng.train[:4]
    x, xm, yg, ygm, yl, ylm, ym, ymm = de.prepare_data(eng, ids)
    assert x.dtype == np.int64 and xm.dtype == np.float32 and x.shape == xm.shape
    assert yg.shape == (4, 5, 16) and yl.shape == (4, 5, 3, 24) and ym.shape == (4, 5, 20)
    assert ygm.shape == (4, 5) and ylm.shape == (4, 5, 3) and ymm.shape == (4, 5)
    for i, ID in enumerate(ids):
        vid, cap = ID.split('_')
        words = [c for c in eng.CAP[vid] if c['cap_id'] == cap][0]['tokenized'].split(' ')
        want = [eng.worddict[w] if eng.worddict[w] < eng.n_words else 1 for w in words]
        n = len(want)
        assert x[:n, i].tolist() == want and (x[n:, i] == 0).all()          # 0 = eos padding
        assert xm[:n + 1, i].tolist() == [1.0] * (n + 1) and (xm[n + 1:, i] == 0).all()
        assert (ygm[i] == (yg[i].sum(-1) != 0)).all()
    assert x.shape[0] == max(int(xm[:, i].sum()) for i in range(4))          # longest caption + eos
    assert (x < eng.n_words).all()                                            # ids beyond n_words -> UNK (1)
    # captions of length >= maxlen are dropped; nothing left -> five Nones
    eng.maxlen = 2
    assert de.prepare_data(eng, ids) == (None, None, None, None, None)
    eng.maxlen = 6
    r = de.prepare_data(eng, eng.train)
    assert r[0].shape[0] <= 6 and r[0].shape[1] == sum(
        len([c for c in eng.CAP[i.split('_')[0]] if c['cap_id'] == i.split('_')[1]][0]['tokenized'].split(' ')) < 6
        for i in eng.train)
    assert [len(b) for b in eng.kf_train] == [4] * (len(eng.train) // 4) + ([len(eng.train) % 4] if len(eng.train) % 4 else [])


def test_pred_probs_with_an_engine_that_has_no_prepare_data_method():
    """The reference's Movie2Caption engine has no prepare_data method (it is a module function,
    data_engine.py:258): pred_probs -- called by train() with the reference's 3-argument signature --
    must fall back to the module function."""
    import video_description_with_spatial_temporal_attention_b200 as stat
    from video_description_with_spatial_temporal_attention_b200 import data_engine as de, model_attention as ma
    o = stat.default_options(dim=16, dim_word=16, ctxg_dim=16, ctxl_dim=24, ctxm_dim=20, n_words=30)
    real = de.synthetic_engine(o, n_videos=6, caps_per_video=2, T=5, R=3, seed=3)

    class Movie2CaptionLike(object):          # attribute access only, like the reference's engine object
        def __init__(self, inner):
            self.__dict__.update({k: v for k, v in inner.__dict__.items()})
            self.signature = inner.signature
            for m in ('get_video_global_features', 'get_video_local_features', 'get_video_motion_features',
                      'get_ctxg_mask', 'get_ctxl_mask', 'get_ctxm_mask'):
                setattr(self, m, getattr(inner, m))
    eng = Movie2CaptionLike(real)
    assert not hasattr(eng, 'prepare_data')
    model = ma.Attention()
    model.engine = eng
    seen = []

    def f_log_probs(*batch):
        assert len(batch) == 8
        seen.append(batch[0].shape)
        return -np.ones(batch[0].shape[1], 'float32')
    err, perp = model.pred_probs('valid', f_log_probs, verbose=False)
    assert seen and err == 1.0 and perp > 1.0


def test_use_dropout_false_is_rejected():
    import video_description_with_spatial_temporal_attention_b200 as stat
    from video_description_with_spatial_temporal_attention_b200 import model_attention as ma
    o = stat.default_options(dim=16, dim_word=16, ctxg_dim=16, ctxl_dim=24, ctxm_dim=20, n_words=30)
    o['use_dropout'] = False
    with pytest.raises(ValueError):
        ma.validate_options(o)


def test_checkpoint_roundtrip(tmp_path):
    import video_description_with_spatial_temporal_attention_b200 as stat
    from video_description_with_spatial_temporal_attention_b200 import checkpoint as ck, model_attention as ma
    o = stat.default_options(dim=8, dim_word=8, ctxg_dim=8, ctxl_dim=12, ctxm_dim=10, n_words=11)
    model = ma.Attention()
    params = model.init_params(o)
    tp = model.init_tparams(params)
    path = str(tmp_path / 'model_best_so_far.npz')
    ck.save_params(path, tp, history_errs=[[1.0, 2.0, 3.0]], train_err=0.5, valid_err=0.6, test_err=0.7)
    ck.save_options(str(tmp_path), o)
    o2 = ck.load_options(str(tmp_path))
    assert o2 == dict(o)
    fresh = model.init_params(o2)
    got = ck.load_params(path, fresh)
    assert list(got.keys()) == list(params.keys())
    for k in params:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], params[k])
    ex = ck.archive_extras(path)
    assert float(ex['valid_err']) == pytest.approx(0.6) and ex['history_errs'].shape == (1, 3)
    # the reference's own loader contract: a missing key / a wrong shape is an error
    bad = dict(params)
    bad.pop('Wemb')
    np.savez(str(tmp_path / 'bad.npz'), **bad)
    with pytest.raises(Exception):
        ck.load_params(str(tmp_path / 'bad.npz'), model.init_params(o))
    o3 = dict(o, n_words=12)
    with pytest.raises(ValueError):
        ck.load_params(path, model.init_params(o3))
    # model.load_params (model_attention.py:1109-1113) reads the same archive
    again = model.load_params(path, model.init_params(o))
    assert np.array_equal(again['decoder_U'], params['decoder_U'])


def test_stochastic_branch_of_gen_sample_and_token_draws():
    """gen_sample(stochastic=True) follows f_next's own next_sample (model_attention.py:914-918): with the
    oracle callables (whose next_sample is the mode) it walks the greedy path and sums PROBABILITIES."""
    g = Golden(NAMES[0])
    x, m, ctxg, mg, ctxl, ml, ctxm, mm = g.batch
    f_init, f_next = so.make_sampler(g.params, g.options, hoist=True)
    model = ma.Attention()
    sample, score, _, _ = model.gen_sample(None, f_init, f_next, ctxg[0], mg[0], ctxl[0], ml[0], ctxm[0], mm[0],
                                           g.options, None, 1, g.maxlen, True)
    want, _ = g.hyps(1, 0)
    assert sample == want[0]
    # replay the path by hand: score = sum_t p_t[w_t]
    r = f_init(ctxg[0], mg[0])
    h, c, w, tot = r[1][None], r[2][None], -np.ones((1,), 'int64'), 0.0
    for tok in sample:
        p, _, h, c = f_next(w, r[0], mg[0], ctxl[0], ml[0], ctxm[0], mm[0], h, c)
        tot += p[0, tok]
        w = np.array([tok], 'int64')
    assert abs(score - tot) < 1e-6
    # the draw helper: right distribution, valid ids, reproducible
    p = np.array([[0.1, 0.2, 0.7], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], 'float32')
    d = np.stack([ma.draw_tokens(p, np.random.RandomState(s)) for s in range(4000)])
    assert (d[:, 1] == 0).all() and (d[:, 2] == 2).all()
    freq = np.bincount(d[:, 0], minlength=3) / 4000.0
    assert np.abs(freq - p[0]).max() < 0.03
    assert np.array_equal(ma.draw_tokens(p, np.random.RandomState(5)), ma.draw_tokens(p, np.random.RandomState(5)))
