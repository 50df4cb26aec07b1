"""Bookkeeping of the training loop (train_loop.train, reference model_attention.py:1211-1558) with scripted
callables in place of the device: update / validation cadence, the history row, best-model selection on the
validation error (including the reference's "nothing is best at the first validation" rule), patience, the
files train() leaves behind."""
import os
from collections import OrderedDict

import numpy as np
import pytest

import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import checkpoint, data_engine, train_loop


class FakeModel(object):
    def __init__(self, engine, valid_errs):
        self.engine = engine
        self.valid_errs = list(valid_errs)
        self.calls = []

    def pred_probs(self, whichset, f_log_probs, verbose=True):
        self.calls.append(whichset)
        if whichset == 'valid':
            e = self.valid_errs.pop(0) if self.valid_errs else 9.0
            return e, 2.0 ** e
        return (1.0, 2.0) if whichset == 'train' else (7.0, 128.0)


class Harness(object):
    def __init__(self, tmp_path, valid_errs, costs=None):
        self.o = stat.default_options(dim=8, dim_word=8, ctxg_dim=8, ctxl_dim=12, ctxm_dim=16, n_words=30)
        self.eng = data_engine.synthetic_engine(self.o, n_videos=8, caps_per_video=3, T=4, R=2, mb_size_train=4,
                                                mb_size_test=5)
        self.model = FakeModel(self.eng, valid_errs)
        self.params = OrderedDict(w=np.zeros(3, 'float32'), b=np.float32(0.).reshape(()))
        self.updates = 0
        self.costs = costs
        self.dir = str(tmp_path) + os.sep
        self.noise = []
        self.shapes = []

    # the callables train() receives
    def f_grad_shared(self, *batch):
        assert len(batch) == 8
        self.shapes.append(batch[0].shape)
        c = self.costs[self.updates] if self.costs else 10.0 / (1 + self.updates)
        return [c, None, None, None, None, None]

    def f_update(self, lr):
        self.updates += 1
        self.params['w'] = np.full(3, self.updates, 'float32')

    def f_alphas(self, *batch):
        B = batch[0].shape[1]
        a = np.tile(np.array([0.1, 0.2, 0.3, 0.4], 'float32'), (2, B, 1))
        return [np.tile(a[..., None], (1, 1, 1, 2)) / 2, a, a, a]

    def get_params(self):
        return OrderedDict((k, np.array(v)) for k, v in self.params.items())

    def set_params(self, p):
        self.params = OrderedDict((k, np.array(v)) for k, v in p.items())

    class Noise(object):
        def __init__(self, log):
            self.log = log

        def set_value(self, v):
            self.log.append(v)

    def run(self, **kw):
        args = dict(lrate=0.01, patience=2, max_epochs=50, dispFreq=3, validFreq=2, sampleFreq=1000)
        args.update(kw)
        return train_loop.train(self.model, data_engine, self.f_grad_shared, self.f_update, lambda *b: None,
                                self.f_alphas, self.get_params, self.set_params, self.o, self.dir,
                                use_noise=self.Noise(self.noise), log=lambda *a: None, **args)


def test_validation_cadence_best_model_and_early_stop(tmp_path):
    # validation errors in order: the 1st can never be "best" (len(history) > 1 rule), the 2nd improves, the 4th
    # improves again, then three non-improvements exceed patience = 2
    errs = [5.0, 4.0, 4.5, 3.0, 3.5, 3.6, 3.7, 1.0, 1.0]
    h = Harness(tmp_path, errs + [3.0])          # the last entry feeds the closing pred_probs('valid')
    tr, va, te = h.run()
    hist = np.loadtxt(h.dir + 'train_valid_test.txt', ndmin=2)
    assert hist.shape == (7, len(train_loop.HISTORY_COLUMNS))            # stopped at the 7th validation
    assert h.updates == 14                                               # validFreq = 2
    np.testing.assert_allclose(hist[:, train_loop.COL_VALID_ERR], errs[:7])
    np.testing.assert_allclose(hist[:, 1], np.arange(2, 15, 2))          # uidx column
    np.testing.assert_allclose(hist[:, 4], 2.0 ** np.array(errs[:7]), rtol=1e-4)   # valid_perp column
    assert (hist[:, 2] == 1.0).all() and (hist[:, 7] == 7.0).all()       # train_err, test_err
    # best parameters: those of the 4th validation (update 8), restored at the end and written out
    assert (h.params['w'] == 8).all()
    best = np.load(h.dir + 'model_best_so_far.npz')
    assert (best['w'] == 8).all() and best['history_errs'].shape == (4, 22)
    final = np.load(h.dir + 'model_best.npz')
    assert (final['w'] == 8).all() and float(final['valid_err']) == 1.0 and final['history_errs'].shape == (7, 22)
    cur = np.load(h.dir + 'model_current.npz')
    assert (cur['w'] == 14).all()
    assert checkpoint.load_options(h.dir)['dim'] == 8
    for k in ('alphal', 'alphag', 'alpham', 'alphalt'):
        r = np.loadtxt(h.dir + k + '_ratio.txt', ndmin=1)
        assert r.shape == (7,)
        # min / max of the scripted weights (the scripted spatial weights are uniform over the two regions)
        np.testing.assert_allclose(r, 1.0 if k == 'alphal' else 0.25, rtol=1e-6)
    assert (tr, va, te) == (1.0, 1.0, 0)
    assert h.noise[0] == 1.0 and h.noise[-1] == 0.0 and 0.0 in h.noise[1:-1]
    # every batch is prepare_data's 8-tuple of this engine
    assert all(s[1] in (4, 3, 2, 1) for s in h.shapes)


def test_first_validation_never_saves_a_best_model(tmp_path):
    h = Harness(tmp_path, [1.0, 2.0, 2.0, 2.0, 2.0])
    h.run(patience=1)
    # the best (first) validation happened before anything could be recorded as best: the initial parameters
    # are what the reference would restore and save, too (:1231, :1472)
    assert not os.path.exists(h.dir + 'model_best_so_far.npz')
    assert (h.params['w'] == 0).all()
    assert (np.load(h.dir + 'model_best.npz')['w'] == 0).all()
    hist = np.loadtxt(h.dir + 'train_valid_test.txt', ndmin=2)
    assert hist.shape[0] == 3                                            # bad_counter 1, 2 > patience -> stop


def test_debug_mode_runs_one_update(tmp_path):
    h = Harness(tmp_path, [])
    tr, va, te = h.run(debug=True, validFreq=1)
    assert h.updates == 1 and h.model.calls == []
    assert (tr, va, te) == (-1, 0, 0)
    hist = np.loadtxt(h.dir + 'train_valid_test.txt', ndmin=2)
    assert hist.shape == (1, 22) and hist[0, train_loop.COL_VALID_ERR] == -1


def test_nan_cost_raises_and_scores_are_recorded(tmp_path):
    h = Harness(tmp_path, [], costs=[1.0, float('nan')])
    with pytest.raises(FloatingPointError):
        h.run()
    assert h.updates == 1
    h2 = Harness(tmp_path, [3.0, 2.0, 2.5])
    scores = {'valid': dict((k, 10.0 + i) for i, k in enumerate(train_loop.SCORE_KEYS)),
              'test': dict((k, 20.0 + i) for i, k in enumerate(train_loop.SCORE_KEYS))}
    h2.run(max_epochs=1, validFreq=3, score_fn=lambda p: scores)
    hist = np.loadtxt(h2.dir + 'train_valid_test.txt', ndmin=2)
    # column order of :1455-1461: B1..B4, meteor, Rouge, Cider
    np.testing.assert_allclose(hist[0, 8:15], [10, 11, 12, 13, 14, 15, 16])
    np.testing.assert_allclose(hist[0, 15:22], [20, 21, 22, 23, 24, 25, 26])
    assert train_loop.alpha_ratio(np.array([[0.5, 0.5], [0.25, 0.75]])) == pytest.approx(0.75 / 1.25)


def test_fit_keeps_the_shared_parameters_in_step_with_the_trainer(tmp_path):
    """fit(): the tparams that f_log_probs reads hold the trainer's current values at every validation and the
    best ones at the end."""
    from video_description_with_spatial_temporal_attention_b200 import model_attention as ma
    h = Harness(tmp_path, [5.0, 4.0, 4.5, 4.6, 4.7, 4.8, 4.0])
    tparams = ma.Attention().init_tparams(h.get_params())
    seen = []

    class FakeTrainer(object):
        f_grad_shared = staticmethod(h.f_grad_shared)
        f_update = staticmethod(h.f_update)
        unzip = staticmethod(h.get_params)
        load_params = staticmethod(h.set_params)

    def f_log_probs(*batch):
        return None

    real_pred = h.model.pred_probs

    def pred_probs(whichset, f, verbose=True):
        if whichset == 'valid':
            seen.append(float(tparams['w'].get_value()[0]))
        return real_pred(whichset, f, verbose)

    h.model.pred_probs = pred_probs
    train_loop.fit(h.model, tparams, h.o, FakeTrainer, f_log_probs, h.f_alphas, h.dir, log=lambda *a: None,
                   patience=2, max_epochs=50, validFreq=2, sampleFreq=1000)
    # validations after updates 2, 4, ... saw exactly those parameter values; the closing one the best (update 4)
    assert seen == [2.0, 4.0, 6.0, 8.0, 10.0, 4.0]
    assert float(tparams['w'].get_value()[0]) == 4.0 and (h.params['w'] == 4).all()


def test_reload_continues_the_history(tmp_path):
    """reload_ (:1109-1113, :1215-1218): parameters and history of model_best_so_far.npz feed the next run, whose
    validations are judged against the reloaded history."""
    h = Harness(tmp_path, [5.0, 4.0, 4.5, 4.6, 4.7, 4.0])
    h.run()
    params, hist = train_loop.reload(h.dir, OrderedDict(w=np.zeros(3, 'float32'), b=np.zeros((), 'float32')))
    assert (params['w'] == 4).all() and len(hist) == 2 and hist[-1][train_loop.COL_VALID_ERR] == 4.0
    h2 = Harness(tmp_path, [3.9, 5.0, 5.0, 5.0, 3.9])
    h2.set_params(params)
    h2.run(history_errs=hist, patience=1)
    out = np.loadtxt(h2.dir + 'train_valid_test.txt', ndmin=2)
    # two reloaded rows, then 3.9 is an improvement at once (the history is longer than one row), then two bad ones
    assert out.shape[0] == 2 + 3 and (np.load(h2.dir + 'model_best_so_far.npz')['w'] == 2).all()
