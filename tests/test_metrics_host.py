"""Caption generation for scoring (metrics.py mirror of reference metrics.py:103-182) with a scripted search in
place of the device: which clips are decoded, which hypothesis is kept, how ids become words, the files and the
sample-pair dicts; per-clip (gen_sample) and batched (beam_batch) routes give the same result."""
import numpy as np

import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import data_engine, metrics


class ScriptedSearch(object):
    """hypotheses depend on the clip's features only, so both routes must agree"""

    def __init__(self):
        self.calls = []

    def _hyps(self, ctxg):
        k = int(abs(float(np.asarray(ctxg).sum())) * 1000) % 20 + 2
        return [[k, k + 1, 0], [k + 2, 999999, 0, 5]], [2.0, 1.0]      # the second one is cheaper; 999999 -> UNK

    def gen_sample(self, tparams, f_init, f_next, ctxg, mg, ctxl, ml, ctxm, mm, options, trng=None, k=1, maxlen=30,
                   stochastic=False):
        assert tparams is None and k == 5 and maxlen == metrics.MAXLEN
        assert mg.shape == ctxg.shape[:1] and ml.shape == ctxl.shape[:2]
        self.calls.append('clip')
        h, s = self._hyps(ctxg)
        return h, s, None, None

    def beam_batch(self, tparams, options, ctxg, mg, ctxl, ctxm, k=5, maxlen=30, use_graph=False):
        assert ctxg.ndim == 3 and ctxl.ndim == 4 and maxlen == metrics.MAXLEN
        self.calls.append('batch%d' % ctxg.shape[0])
        return [self._hyps(ctxg[b]) for b in range(ctxg.shape[0])]


class Scorer(object):
    def score(self, gts, samples, ids):
        assert list(gts.keys()) == list(ids) == list(samples.keys())
        return dict((k, float(len(ids))) for k in metrics.SCORE_KEYS)


def test_samples_files_pairs_and_both_routes(tmp_path):
    o = stat.default_options(dim=8, dim_word=8, ctxg_dim=8, ctxl_dim=12, ctxm_dim=16, n_words=30)
    eng = data_engine.synthetic_engine(o, n_videos=8, caps_per_video=3, T=4, R=2)
    assert eng.valid_ids == ['vid5', 'vid6'] and eng.test_ids == ['vid7', 'vid8'] and len(eng.train_ids) == 4
    g, gm, l, lm, m, mm = eng.prepare_data_for_blue('valid')
    assert len(g) == 2 and g[0].shape == (4, 8) and lm[0].shape == (4, 2) and mm[0].shape == (4,)
    model = ScriptedSearch()
    d1, d2 = str(tmp_path / 'a'), str(tmp_path / 'b')
    import os
    os.makedirs(d1), os.makedirs(d2)
    sv, st = metrics.generate_sample_gpu_single_process('attention', None, o, eng, model, None, None, save_dir=d1,
                                                        beam=5, whichset='both')
    assert model.calls == ['clip'] * 4
    model.calls = []
    bv, bt = metrics.generate_sample_gpu_single_process('attention', None, o, eng, model, None, None, save_dir=d2,
                                                        beam=5, whichset='both', tparams=object(), batch_size=3)
    assert model.calls == ['batch2', 'batch2']
    assert sv == bv and st == bt
    assert list(sv.keys()) == eng.valid_ids and sv['vid5'][0]['image_id'] == 'vid5'
    for d in (d1, d2):
        lines = open(os.path.join(d, 'valid_samples.txt')).read().split('\n')
        assert lines[:2] == [sv[v][0]['caption'] for v in eng.valid_ids]
    # the cheaper hypothesis, cut at the first 0, out-of-dictionary ids as UNK
    cap = sv['vid5'][0]['caption'].split(' ')
    assert len(cap) == 2 and cap[1] == 'UNK' and cap[0] == eng.word_idict[model._hyps(g[0])[0][1][0]]
    # only one split asked for
    only_v, none_t = metrics.generate_sample_gpu_single_process('attention', None, o, eng, model, None, None,
                                                                save_dir=d1, whichset='valid')
    assert none_t is None and only_v == sv


def test_compute_score_contract(tmp_path):
    o = stat.default_options(dim=8, dim_word=8, ctxg_dim=8, ctxl_dim=12, ctxm_dim=16, n_words=30)
    eng = data_engine.synthetic_engine(o, n_videos=8, caps_per_video=3, T=4, R=2)
    model = ScriptedSearch()
    res = metrics.compute_score('attention', None, o, eng, str(tmp_path), 5, 5, 'both', False, None, None, None, None,
                                False, 'blue', None, None, model, scorer=Scorer())
    scores, processes, queue, rqueue, shared = res
    assert scores['valid']['Bleu_4'] == 2.0 and scores['test']['CIDEr'] == 2.0 and processes is None
    one = metrics.compute_score('attention', None, o, eng, str(tmp_path), 5, 5, whichset='valid', on_cpu=False,
                                one_time=True, model=model)
    assert one['test'] is None and set(one['valid']) == set(metrics.SCORE_KEYS)
    import pytest
    with pytest.raises(NotImplementedError):
        metrics.compute_score('attention', None, o, eng, str(tmp_path), 5, 5, model=model)
