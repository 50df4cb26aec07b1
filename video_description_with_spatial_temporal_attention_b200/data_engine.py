"""Minibatch side of the path: the 8-tuple ``prepare_data`` hands to ``f_log_probs`` /
``f_grad_shared`` (reference ``data_engine.py:258-337``) and an in-memory engine with the
attributes ``train()`` and ``pred_probs`` touch (``data_engine.py:20-60,118-255``).

The reference engine reads MSVD features from HDF5 / pickles (out of scope); this
one is fed numpy arrays, keeps the same attribute and method names, and produces batches that
are bit-identical in layout: ``x`` (L,B) int64 padded with 0, ``mask`` (L,B) with one extra 1
for the end-of-sentence position, features stacked per caption, masks = "row is not all zero".
"""
from __future__ import annotations

from collections import OrderedDict

import numpy

from . import common


def _nonzero_rows(feat, dim):
    # data_engine.py:169-219: a frame / region counts when its first `dim` features do not sum to 0
    feat = numpy.asarray(feat)
    return (feat[..., :dim].sum(axis=-1) != 0).astype('int32').astype('float32')


class MemoryEngine(object):
    """signature 'youtube2text' ids: '<vidID>_<capID>' (data_engine.py:276-277).

    features: mapping vidID -> (ctxg (T,Dg), ctxl (T,R,Dr), ctxm (T,Dm)) float32
    captions: mapping vidID -> list of {'cap_id': str, 'tokenized': 'w1 w2 ...'}
    worddict: mapping word -> id (ids >= n_words become 1 = UNK, data_engine.py:295-296)
    """
    signature = 'youtube2text'

    def __init__(self, features, captions, worddict, n_words, train, valid, test, mb_size_train=64,
                 mb_size_test=128, maxlen=None):
        self.features = features
        self.CAP = captions
        self.worddict = worddict
        self.n_words = n_words
        self.maxlen = maxlen
        self.mb_size_train, self.mb_size_test = mb_size_train, mb_size_test
        self.train, self.valid, self.test = list(train), list(valid), list(test)
        g, l, m = next(iter(features.values()))
        self.ctxg_dim, self.ctxl_dim, self.ctxm_dim = g.shape[-1], l.shape[-1], m.shape[-1]
        self.word_idict = OrderedDict((v, k) for k, v in worddict.items())
        self.word_idict[0] = '<eos>'
        self.word_idict[1] = 'UNK'
        # clip ids of each split, in order of first appearance (data_engine.py:229-232: vid1..1200 / 1201..1300 / ...)
        uniq = lambda ids: list(OrderedDict((i.rsplit('_', 1)[0], None) for i in ids))
        self.train_ids, self.valid_ids, self.test_ids = uniq(self.train), uniq(self.valid), uniq(self.test)
        self.kf_train = common.generate_minibatch_idx(len(self.train), min(mb_size_train, len(self.train)))
        self.kf_valid = common.generate_minibatch_idx(len(self.valid), min(mb_size_test, len(self.valid)))
        self.kf_test = common.generate_minibatch_idx(len(self.test), min(mb_size_test, len(self.test)))

    def get_video_global_features(self, vid):
        return self.features[vid][0]

    def get_video_local_features(self, vid):
        return self.features[vid][1]

    def get_video_motion_features(self, vid):
        return self.features[vid][2]

    def get_ctxg_mask(self, ctxg):
        return _nonzero_rows(ctxg, self.ctxg_dim)

    def get_ctxl_mask(self, ctxl):
        return _nonzero_rows(ctxl, self.ctxl_dim)

    def get_ctxm_mask(self, ctxm):
        return _nonzero_rows(ctxm, self.ctxm_dim)

    def prepare_data(self, engine, IDs):
        return prepare_data(engine, IDs)

    def prepare_data_for_blue(self, whichset):
        """Features and masks of every clip of a split, one entry per clip (data_engine.py:137-167)."""
        ids = {'valid': self.valid_ids, 'test': self.test_ids, 'train': self.train_ids}[whichset]
        g = [self.get_video_global_features(v) for v in ids]
        l = [self.get_video_local_features(v) for v in ids]
        m = [self.get_video_motion_features(v) for v in ids]
        return (g, [self.get_ctxg_mask(a) for a in g], l, [self.get_ctxl_mask(a) for a in l],
                m, [self.get_ctxm_mask(a) for a in m])


def prepare_data(engine, IDs):
    """-> (x, x_mask, ctxg, ctxg_mask, ctxl, ctxl_mask, ctxm, ctxm_mask), the positional inputs of
    f_log_probs; captions of length >= engine.maxlen are dropped; five Nones when none is left."""
    rows = []
    for ID in IDs:
        if engine.signature == 'youtube2text':
            vid, cap = ID.split('_')
        elif engine.signature == 'lsmdc':
            parts = ID.split('_')
            vid, cap = '_'.join(parts[:-1]), parts[-1]
        else:
            raise NotImplementedError(engine.signature)
        words = None
        for c in engine.CAP[vid]:
            if c['cap_id'] == cap:
                words = c['tokenized'].split(' ')
                break
        assert words is not None, ID
        seq = [engine.worddict[w] if engine.worddict[w] < engine.n_words else 1 for w in words]
        rows.append((seq, engine.get_video_global_features(vid), engine.get_video_local_features(vid),
                     engine.get_video_motion_features(vid)))
    if engine.maxlen is not None:
        rows = [r for r in rows if len(r[0]) < engine.maxlen]
        if len(rows) < 1:
            return None, None, None, None, None
    lengths = [len(r[0]) for r in rows]
    yg = numpy.asarray([r[1] for r in rows])
    yl = numpy.asarray([r[2] for r in rows])
    ym = numpy.asarray([r[3] for r in rows])
    L = int(numpy.max(lengths)) + 1
    x = numpy.zeros((L, len(rows))).astype('int64')
    x_mask = numpy.zeros((L, len(rows))).astype('float32')
    for i, (seq, _, _, _) in enumerate(rows):
        x[:lengths[i], i] = seq
        x_mask[:lengths[i] + 1, i] = 1.
    return (x, x_mask, yg, engine.get_ctxg_mask(yg), yl, engine.get_ctxl_mask(yl), ym,
            engine.get_ctxm_mask(ym))


def synthetic_engine(options, n_videos=12, caps_per_video=3, T=26, R=8, seed=0, maxlen=None,
                     mb_size_train=4, mb_size_test=5):
    """A seeded in-memory engine of MSVD shape (synthetic features and word ids) for tests and
    examples: every clip has caps_per_video captions of 2..9 words; some trailing frames are zero."""
    rng = numpy.random.RandomState(seed)
    V = options['n_words']
    worddict = OrderedDict(('w%d' % i, i) for i in range(2, V + 3))       # a few ids beyond n_words -> UNK
    feats, caps = OrderedDict(), OrderedDict()
    for v in range(1, n_videos + 1):
        vid = 'vid%d' % v
        g = (rng.randn(T, options['ctxg_dim']) * 0.5).astype('float32')
        l = (rng.randn(T, R, options['ctxl_dim']) * 0.5).astype('float32')
        m = (rng.randn(T, options['ctxm_dim']) * 0.5).astype('float32')
        z = rng.randint(0, max(1, T // 4))
        if z:
            g[T - z:] = 0
            l[T - z:] = 0
            m[T - z:] = 0
        feats[vid] = (g, l, m)
        caps[vid] = [{'cap_id': str(c), 'tokenized': ' '.join('w%d' % rng.randint(2, V + 3)
                                                               for _ in range(rng.randint(2, 10)))}
                     for c in range(caps_per_video)]
    ids = ['%s_%d' % (vid, c) for vid in feats for c in range(caps_per_video)]
    n = len(ids)
    return MemoryEngine(feats, caps, worddict, V, ids[:n // 2], ids[n // 2:3 * n // 4], ids[3 * n // 4:],
                        mb_size_train=mb_size_train, mb_size_test=mb_size_test, maxlen=maxlen)
