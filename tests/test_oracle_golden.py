"""The oracle (oracle/stat_oracle.py) against fixtures produced by executing the
reference's own source (tests/golden/make_golden.py).  This is what pins the
oracle; everything GPU-side is then compared with the oracle."""
import numpy as np
import pytest

from oracle import ref_exec, stat_oracle as so
from tests.golden_util import NAMES, Golden


def test_fixtures_present():
    assert len(NAMES) >= 4


@pytest.mark.parametrize('name', NAMES)
def test_f_log_probs_and_intermediates(name):
    g = Golden(name)
    lp, allv = so.forward_teacher(g.params, g.options, *g.batch, return_all=True)
    np.testing.assert_allclose(lp, g.out('f_log_probs'), rtol=0, atol=2e-5)
    st = allv['steps']
    np.testing.assert_allclose(np.stack([s['alphaL'] for s in st]), g.out('alphals'), atol=2e-6)
    np.testing.assert_allclose(np.stack([s['alphaG'] for s in st]), g.out('alphags'), atol=2e-6)
    np.testing.assert_allclose(np.stack([s['alphaM'] for s in st]), g.out('alphams'), atol=2e-6)
    np.testing.assert_allclose(np.stack([s['alphaLT'] for s in st]), g.out('alphalts'), atol=2e-6)
    np.testing.assert_allclose(np.concatenate([s['probs'] for s in st]), g.out('probs'), atol=2e-6)
    # fp64 twin bounds the fp32 oracle's own rounding
    lp64 = so.forward_teacher(g.params, g.options, *g.batch, dtype=np.float64)
    np.testing.assert_allclose(lp64, g.out('f_log_probs'), atol=2e-5)


@pytest.mark.parametrize('name', NAMES)
@pytest.mark.parametrize('hoist', [False, True])
def test_sampler_functions(name, hoist):
    g = Golden(name)
    x, m, ctxg, mg, ctxl, ml, ctxm, mm = g.batch
    f_init, f_next = so.make_sampler(g.params, g.options, hoist=hoist)
    r = f_init(ctxg[0], mg[0])
    np.testing.assert_allclose(r[1], g.out('f_init_h0'), atol=2e-6)
    np.testing.assert_allclose(r[2], g.out('f_init_c0'), atol=2e-6)
    r1 = f_next(-np.ones((1,), 'int64'), ctxg[0], mg[0], ctxl[0], ml[0], ctxm[0], mm[0],
                r[1][None], r[2][None])
    np.testing.assert_allclose(r1[0], g.out('f_next0_probs'), atol=2e-6)
    np.testing.assert_allclose(r1[2], g.out('f_next0_h'), atol=2e-6)
    np.testing.assert_allclose(r1[3], g.out('f_next0_c'), atol=2e-6)
    r2 = f_next(g.inp('f_next1_x'), ctxg[0], mg[0], ctxl[0], ml[0], ctxm[0], mm[0],
                g.inp('f_next1_h'), g.inp('f_next1_c'))
    np.testing.assert_allclose(r2[0], g.out('f_next1_probs'), atol=2e-6)
    np.testing.assert_allclose(r2[2], g.out('f_next1_h'), atol=2e-6)
    np.testing.assert_allclose(r2[3], g.out('f_next1_c'), atol=2e-6)


@pytest.mark.parametrize('name', NAMES)
def test_gen_sample_tokens_and_scores(name):
    g = Golden(name)
    x, m, ctxg, mg, ctxl, ml, ctxm, mm = g.batch
    f_init, f_next = so.make_sampler(g.params, g.options, hoist=True)
    n_dead_early = n_maxlen = 0
    for k in g.ks:
        for b in range(ctxg.shape[0]):
            want, want_sc = g.hyps(k, b)
            got, got_sc, _, _ = so.gen_sample(f_init, f_next, ctxg[b], mg[b], ctxl[b], ml[b],
                                              ctxm[b], mm[b], k=k, maxlen=g.maxlen)
            assert got == want, (name, k, b)
            np.testing.assert_allclose(np.asarray(got_sc), want_sc, atol=2e-5)
            n_dead_early += sum(len(h) < g.maxlen for h in want)
            n_maxlen += sum(len(h) == g.maxlen and h[-1] != 0 for h in want)
    if name.endswith('_init'):
        assert n_maxlen > 0          # a hypothesis that hits maxlen
    else:
        assert n_dead_early > 0      # a hypothesis that retires early (shrinking beam)


@pytest.mark.parametrize('name', NAMES)
def test_greedy_batch_equals_gen_sample_k1(name):
    g = Golden(name)
    x, m, ctxg, mg, ctxl, ml, ctxm, mm = g.batch
    toks, lens, scores = so.greedy_decode_batch(g.params, g.options, ctxg, mg, ctxl, ctxm,
                                                g.maxlen)
    for b in range(ctxg.shape[0]):
        want, want_sc = g.hyps(1, b)
        assert [int(t) for t in toks[b, :lens[b]]] == want[0]
        np.testing.assert_allclose(scores[b], want_sc[0], atol=2e-5)


def test_init_params_matches_reference_layout():
    """Key order / shapes of App. B; and, where the reference is mounted, bit
    equality with the reference's own init_params."""
    o = so.default_options(dim_word=8, dim=8, ctxg_dim=8, ctxl_dim=12, ctxm_dim=10, n_words=11)
    p = so.init_params(o)
    assert len(p) == 41
    assert list(p)[:5] == ['Wemb', 'ff_state_W', 'ff_state_b', 'ff_memory_W', 'ff_memory_b']
    assert list(p)[-2:] == ['ff_logit_W', 'ff_logit_b']
    assert p['decoder_U'].shape == (8, 32) and p['decoder_b_sel'].shape == ()
    g = Golden('ref_tiny_init')
    assert list(g.params) == list(p)
    for k in p:
        np.testing.assert_array_equal(np.asarray(p[k]), g.params[k])
    o2 = so.default_options(dim=8, dim_word=8, ctxg_dim=20, global_proj=True, n_words=11,
                            ctxl_dim=12, ctxm_dim=10)
    p2 = so.init_params(o2)
    assert len(p2) == 43 and list(p2)[5:7] == ['ff_global_W', 'ff_global_b']
    # D1: the reference's own init_params with its ff_global lines un-commented (ref_exec) draws the
    # same 43 tensors in the same order, bit for bit
    g2 = Golden('ref_tiny_globalproj_init')
    assert list(g2.params) == list(p2)
    for k in p2:
        np.testing.assert_array_equal(np.asarray(p2[k]), g2.params[k])


@pytest.mark.skipif(not ref_exec.available(),
                    reason='opt-in (STAT_RUN_REFERENCE=1 with /root/reference mounted): executes reference source')
@pytest.mark.parametrize('name', ['ref_tiny_trained', 'ref_tiny_globalproj'])
def test_live_reference_execution_matches_fixture(name):
    """In the authoring container: re-run the reference source and check the
    committed fixture is what it produces (fixtures are not stale)."""
    g = Golden(name)
    rm = ref_exec.RefModel(g.options, params=g.params)
    np.testing.assert_allclose(rm.f_log_probs(*g.batch), g.out('f_log_probs'), atol=1e-6)


def test_properties():
    """alpha rows sum to 1; f_next chained == teacher-forced forward."""
    g = Golden('ref_mid_trained')
    x, m, ctxg, mg, ctxl, ml, ctxm, mm = g.batch
    lp, allv = so.forward_teacher(g.params, g.options, *g.batch, return_all=True)
    for s in allv['steps']:
        for a in ('alphaL', 'alphaG', 'alphaM', 'alphaLT'):
            np.testing.assert_allclose(s[a].sum(-1), 1.0, atol=1e-5)
    f_init, f_next = so.make_sampler(g.params, g.options)
    b = 1
    r = f_init(ctxg[b], mg[b])
    h, c = r[1][None], r[2][None]
    prev = -np.ones((1,), 'int64')
    acc = 0.0
    for t in range(x.shape[0]):
        p, _, h2, c2 = f_next(prev, ctxg[b], mg[b], ctxl[b], ml[b], ctxm[b], mm[b], h, c)
        if m[t, b] > 0:
            acc += np.log(p[0, x[t, b]] + 1e-8)
            h, c = h2, c2
        prev = x[t:t + 1, b]
    np.testing.assert_allclose(acc, lp[b], atol=2e-5)
