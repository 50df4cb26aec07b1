"""Caption generation for scoring: the consumer of gen_sample on the validation / test splits
(reference ``metrics.py:103-182``): decode every clip of a split with beam search, keep the cheapest
hypothesis, turn ids into words, write ``valid_samples.txt`` / ``test_samples.txt``, hand
``{vidID: [{'image_id', 'caption'}]}`` to the COCO caption scorer.  Host logic; the search itself is the
device path (``Attention.gen_sample`` over f_init / f_next, or ``Attention.beam_batch`` = the same search for
many clips in one device pass -- same hypotheses in the same order, tests/test_gpu_parity.py).
The scorer (pycocoevalcap: BLEU / METEOR / ROUGE_L / CIDEr, Java) is not part of this repository: pass any object
with ``score(gts, samples, ids) -> dict``; without one ``compute_score`` returns zeros for the metric keys.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy

MAXLEN = 50                                    # metrics.py:15
SCORE_KEYS = ('Bleu_1', 'Bleu_2', 'Bleu_3', 'Bleu_4', 'METEOR', 'ROUGE_L', 'CIDEr')


def seqs2words(caps, word_idict):
    """ids -> 'w1 w2 ...' up to the first 0; ids beyond the dictionary size read as UNK (metrics.py:109-119,
    including its `w > len(word_idict)` test)."""
    out = []
    for cc in caps:
        ww = []
        for w in cc:
            if w == 0:
                break
            ww.append(word_idict[1] if w > len(word_idict) else word_idict[w])
        out.append(' '.join(ww))
    return out


def build_sample_pairs(samples, vidIDs):
    """metrics.py:79-83"""
    D = OrderedDict()
    for sample, vidID in zip(samples, vidIDs):
        D[vidID] = [{'image_id': vidID, 'caption': sample}]
    return D


def generate_sample_gpu_single_process(model_type, model_archive, options, engine, model, f_init, f_next,
                                       save_dir='./samples', beam=5, whichset='both', tparams=None, batch_size=32):
    """metrics.py:103-152.  With `tparams` the clips are searched `batch_size` at a time on the device
    (``model.beam_batch``); without, one clip at a time through ``model.gen_sample(None, f_init, f_next, ...)`` as
    the reference does.  Returns (samples_valid, samples_test) as sample-pair dicts (None for a split not asked for)."""

    def sample(which):
        ctxgs, ctxg_masks, ctxls, ctxl_masks, ctxms, ctxm_masks = engine.prepare_data_for_blue(which)
        best = []
        if tparams is not None:
            for i in range(0, len(ctxgs), batch_size):
                sl = slice(i, i + batch_size)
                res = model.beam_batch(tparams, options, numpy.asarray(ctxgs[sl]), numpy.asarray(ctxg_masks[sl]),
                                       numpy.asarray(ctxls[sl]), numpy.asarray(ctxms[sl]), k=beam, maxlen=MAXLEN)
                for hyps, scores in res:
                    best.append(hyps[int(numpy.argmin(scores))])
        else:
            for ctxg, ctxg_mask, ctxl, ctxl_mask, ctxm, ctxm_mask in zip(ctxgs, ctxg_masks, ctxls, ctxl_masks, ctxms,
                                                                         ctxm_masks):
                hyps, scores, _, _ = model.gen_sample(None, f_init, f_next, ctxg, ctxg_mask, ctxl, ctxl_mask, ctxm,
                                                      ctxm_mask, options, None, beam, maxlen=MAXLEN)
                best.append(hyps[int(numpy.argmin(scores))])
        return seqs2words(best, engine.word_idict)

    samples_valid = samples_test = None
    if whichset in ('valid', 'both'):
        words = sample('valid')
        with open(os.path.join(save_dir, 'valid_samples.txt'), 'w') as f:
            f.write('\n'.join(words) + '\n')
        samples_valid = build_sample_pairs(words, engine.valid_ids)
    if whichset in ('test', 'both'):
        words = sample('test')
        with open(os.path.join(save_dir, 'test_samples.txt'), 'w') as f:
            f.write('\n'.join(words) + '\n')
        samples_test = build_sample_pairs(words, engine.test_ids)
    return samples_valid, samples_test


def score_with_cocoeval(samples_valid, samples_test, engine, scorer):
    """metrics.py:85-101 with the scorer passed in."""
    def one(samples, ids):
        if not samples:
            return None
        gts = OrderedDict((vid, engine.CAP[vid]) for vid in ids)
        return scorer.score(gts, samples, ids)
    return one(samples_valid, engine.valid_ids), one(samples_test, engine.test_ids)


def compute_score(model_type, model_archive, options, engine, save_dir, beam, n_process, whichset='both',
                  on_cpu=True, processes=None, queue=None, rqueue=None, shared_params=None, one_time=False,
                  metric=None, f_init=None, f_next=None, model=None, scorer=None, tparams=None):
    """metrics.py:154-182 (same positional signature; `scorer`, `tparams` are additions)."""
    assert metric != 'perplexity'
    if on_cpu:
        raise NotImplementedError()
    assert model is not None
    samples_valid, samples_test = generate_sample_gpu_single_process(
        model_type, model_archive, options, engine, model, f_init, f_next, save_dir=save_dir, beam=beam,
        whichset=whichset, tparams=tparams)
    if scorer is not None:
        valid_score, test_score = score_with_cocoeval(samples_valid, samples_test, engine, scorer)
    else:
        zeros = dict((k, 0.) for k in SCORE_KEYS)
        valid_score = dict(zeros) if samples_valid else None
        test_score = dict(zeros) if samples_test else None
    scores_final = {'valid': valid_score, 'test': test_score}
    if one_time:
        return scores_final
    return scores_final, processes, queue, rqueue, shared_params
