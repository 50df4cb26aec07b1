"""The bookkeeping of the reference's ``Attention.train`` loop (model_attention.py:1211-1558), around
the device callables: epochs over ``engine.kf_train``, the running train cost, the validation block every
``validFreq`` updates (attention min/max ratios, ``model_current.npz``, errors and perplexities from
``pred_probs``, caption scores, the 22-column ``history_errs`` row -> ``train_valid_test.txt``), best-model
selection on the validation error with ``patience`` early stopping, and the closing ``model_best.npz``.

Host logic only.  Everything that computes is passed in:
  f_grad_shared(*batch) -> [cost, probs, alphals, alphags, alphams, alphalts, ...]    (train.Trainer.f_grad_shared)
  f_update(lrate)                                                                      (train.Trainer.f_update)
  f_log_probs, f_alphas                                                                (model_attention.function)
  get_params() -> OrderedDict name -> ndarray       what ``unzip(tparams)`` is in the reference
  set_params(params)                                what ``zipp(best_p, tparams)`` is
  score_fn(params) -> {'valid': {...}, 'test': {...}} with the keys Bleu_1..4, METEOR, ROUGE_L, CIDEr: the
      reference calls metrics.compute_score (COCO caption metrics, Java; out of scope) -- None gives zeros.
Differences from the reference, all deliberate: a NaN / inf cost raises FloatingPointError instead of entering
pdb (:1270-1275); the motion-attention ratio log stores its own ratio (the reference appends the global one
again, :1378-1379); ``sampleFreq`` printing of decoded samples (:1310-1362) is left to the caller
(``on_sample``).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy

from . import checkpoint, common

SCORE_KEYS = ('Bleu_1', 'Bleu_2', 'Bleu_3', 'Bleu_4', 'METEOR', 'ROUGE_L', 'CIDEr')
# history_errs columns (:1455-1461)
HISTORY_COLUMNS = ('eidx', 'uidx', 'train_err', 'train_perp', 'valid_perp', 'test_perp', 'valid_err', 'test_err',
                   'valid_B1', 'valid_B2', 'valid_B3', 'valid_B4', 'valid_meteor', 'valid_Rouge', 'valid_Cider',
                   'test_B1', 'test_B2', 'test_B3', 'test_B4', 'test_meteor', 'test_Rouge', 'test_Cider')
COL_VALID_ERR, COL_VALID_B4 = 6, 11


def alpha_ratio(alphas):
    """`alphas.min(-1).mean() / alphas.max(-1).mean()` (:1366-1385): 1 = uniform attention, 0 = peaked."""
    a = numpy.asarray(alphas)
    return float(a.min(-1).mean() / a.max(-1).mean())


def reload(from_dir, params):
    """`reload_=True` (:1109-1113, :1215-1218): parameters and the error history of `from_dir/model_best_so_far.npz`.
    Returns (params filled from the archive, history_errs as a list of rows) -- pass the latter to train()."""
    path = os.path.join(from_dir, 'model_best_so_far.npz')
    params = checkpoint.load_params(path, params)
    extras = checkpoint.archive_extras(path)
    hist = extras.get('history_errs')
    return params, (numpy.asarray(hist).tolist() if hist is not None and numpy.size(hist) else [])


def _zero_scores():
    return {s: dict((k, 0.) for k in SCORE_KEYS) for s in ('valid', 'test')}


def train(model, data_engine_module, f_grad_shared, f_update, f_log_probs, f_alphas, get_params, set_params,
          model_options, save_model_dir, use_noise=None, lrate=0.01, patience=10, max_epochs=5000, dispFreq=100,
          validFreq=10, sampleFreq=10, debug=False, verbose=False, score_fn=None, on_sample=None,
          history_errs=None, log=print, grad_norm2=None):
    """Runs the loop; returns (train_err, valid_err, test_err) like the reference (:1558).
    `model` is the Attention host mirror (its `.engine` is the data engine; `pred_probs` is called on it),
    `data_engine_module.prepare_data(engine, tags)` builds the batches (data_engine.py:258-337)."""
    eng = model.engine
    # pred_probs is called with the reference's own 3-argument signature (:996); the host mirror falls back to
    # the module-level prepare_data (data_engine.py:258) when the data engine has no such method, as the
    # reference's Movie2Caption has not
    path = lambda name: os.path.join(save_model_dir, name)
    history_errs = [list(r) for r in history_errs] if history_errs is not None else []     # reload (:1215-1218)
    ratios = OrderedDict((k, []) for k in ('alphal', 'alphag', 'alpham', 'alphalt'))
    bad_counter = 0
    uidx = 0
    estop = False
    best_p = get_params()                                                                   # :1231
    best_valid_err = 999
    train_err = valid_err = test_err = -1
    train_error = 0.
    eidx = 0

    def noise(v):
        if use_noise is not None:
            use_noise.set_value(v)

    for eidx in range(max_epochs):
        n_samples = 0
        train_costs = []
        log('Epoch ', eidx)
        for idx in eng.kf_train:
            tags = [eng.train[i] for i in idx]
            n_samples += len(tags)
            uidx += 1
            noise(1.)
            batch = data_engine_module.prepare_data(eng, tags)
            if batch[0] is None:                                                            # :1253-1255
                log('Minibatch with zero sample under length ', getattr(eng, 'maxlen', None))
                continue
            rvals = f_grad_shared(*batch)
            cost = rvals[0]
            if numpy.isnan(cost) or numpy.isinf(cost):                                      # :1273-1275
                raise FloatingPointError('NaN / inf detected in cost at update %d' % uidx)
            # the reference's grad_nan_report over rvals[6:] (:1263-1269): the squared global gradient norm is
            # already on the device (the clip's first stage); a NaN / inf there must stop the update
            g2 = grad_norm2() if grad_norm2 is not None else None
            if g2 is not None and not numpy.isfinite(g2):
                raise FloatingPointError('NaN / inf detected in the gradients at update %d' % uidx)
            f_update(lrate)
            train_error = cost if eidx == 0 else train_error * 0.95 + cost * 0.05          # :1280-1283
            train_costs.append(cost)

            if numpy.mod(uidx, dispFreq) == 0:
                log('Epoch ', eidx, 'Update ', uidx, 'Train cost mean so far', train_error)

            if numpy.mod(uidx, sampleFreq) == 0 and on_sample is not None:
                noise(0.)
                on_sample(batch)

            if validFreq != -1 and numpy.mod(uidx, validFreq) == 0:
                noise(0.)
                current_params = get_params()          # first: with fit() this also refreshes what f_alphas reads
                al, ag, am, alt = f_alphas(*batch)                                          # :1366-1385
                for k, a in (('alphal', al), ('alphag', ag), ('alpham', am), ('alphalt', alt)):
                    ratios[k].append(alpha_ratio(a))
                    numpy.savetxt(path(k + '_ratio.txt'), ratios[k])
                checkpoint.save_params(path('model_current.npz'), current_params, history_errs)   # :1387-1390
                train_err = train_perp = valid_err = valid_perp = test_err = test_perp = -1
                if not debug:                                                               # :1399-1428
                    train_err, train_perp = model.pred_probs('train', f_log_probs, verbose=verbose)
                    valid_err, valid_perp = model.pred_probs('valid', f_log_probs, verbose=verbose)
                    test_err, test_perp = model.pred_probs('test', f_log_probs, verbose=verbose)
                scores = score_fn(current_params) if score_fn is not None else _zero_scores()
                v, t = scores['valid'], scores['test']
                history_errs.append([eidx, uidx, train_err, train_perp, valid_perp, test_perp, valid_err, test_err,
                                     v['Bleu_1'], v['Bleu_2'], v['Bleu_3'], v['Bleu_4'], v['METEOR'], v['ROUGE_L'],
                                     v['CIDEr'],
                                     t['Bleu_1'], t['Bleu_2'], t['Bleu_3'], t['Bleu_4'], t['METEOR'], t['ROUGE_L'],
                                     t['CIDEr']])
                numpy.savetxt(path('train_valid_test.txt'), history_errs, fmt='%.3f')       # :1462-1463
                hist = numpy.array(history_errs)
                # best caption score: the archive receives best_p, i.e. the best-by-validation-error
                # parameters so far, exactly as the reference does (:1466-1471)
                if len(history_errs) > 1 and v['Bleu_4'] > hist[:-1, COL_VALID_B4].max():
                    checkpoint.save_params(path('model_best_blue_or_meteor.npz'), best_p, history_errs)
                if len(history_errs) > 1 and valid_err < hist[:-1, COL_VALID_ERR].min():    # :1472-1483
                    best_p = get_params()
                    bad_counter = 0
                    best_valid_err = valid_err
                    checkpoint.save_params(path('model_best_so_far.npz'), best_p, history_errs)
                    checkpoint.save_options(save_model_dir, model_options)
                elif len(history_errs) > 1 and valid_err >= hist[:-1, COL_VALID_ERR].min():  # :1484-1493
                    bad_counter += 1
                    log('history best ', hist[:, COL_VALID_ERR].min(), 'bad_counter ', bad_counter, 'patience ',
                        patience)
                    if bad_counter > patience:
                        log('Early Stop!')
                        estop = True
                        break
                log('Train ', train_err, 'Valid ', valid_err, 'Test ', test_err, 'best valid err so far',
                    best_valid_err)
            if debug:
                break
        if estop or debug:
            break
        log('This epoch has seen %d samples, train cost %.2f' % (n_samples, numpy.mean(train_costs)))

    log('Optimization ended.')
    if best_p is not None:
        set_params(best_p)                                                                  # :1522-1523
    noise(0.)
    valid_err = test_err = 0                                                                # :1526-1531
    if not debug:
        valid_err, _ = model.pred_probs('valid', f_log_probs, verbose=verbose)
    numpy.savez(path('model_best.npz'), train_err=train_err, valid_err=valid_err, test_err=test_err,
                history_errs=numpy.asarray(history_errs), **best_p)                         # :1545-1548
    if history_errs != []:
        numpy.savetxt(path('train_valid_test.txt'), numpy.asarray(history_errs), fmt='%.4f')   # :1550-1553
    return train_err, valid_err, test_err


def fit(model, tparams, model_options, trainer, f_log_probs, f_alphas, save_model_dir, use_noise=None,
        data_engine_module=None, caption_scorer=None, decode_samples=False, **kw):
    """train() wired to a device trainer (train.Trainer or anything with f_grad_shared / f_update / unzip /
    load_params): the shared parameters `tparams` -- what f_log_probs, f_init / f_next and the checkpoints read --
    are refreshed from the trainer's device buffers at the top of every validation block and receive the best
    parameters at the end, which is what Theano's in-place updates of the shared variables give the reference
    for free (:1387, :1472-1474, :1522-1523)."""
    if data_engine_module is None:
        from . import data_engine as data_engine_module

    def get_params():
        p = trainer.unzip()
        common.zipp(p, tparams)
        return p

    def set_params(p):
        trainer.load_params(p)
        common.zipp(p, tparams)

    if (caption_scorer is not None or decode_samples) and 'score_fn' not in kw:
        # what the reference does at every validation (:1432-1446): beam-5 captions of the validation and test clips
        # (device search for 32 clips at a time), valid_samples.txt / test_samples.txt, scores from the COCO scorer
        from . import metrics

        def score_fn(params):
            return metrics.compute_score('attention', params, model_options, model.engine, save_model_dir, 5, 5,
                                         whichset='both', on_cpu=False, one_time=True, model=model,
                                         scorer=caption_scorer, tparams=tparams)
        kw['score_fn'] = score_fn
    kw.setdefault('grad_norm2', getattr(trainer, 'grad_norm2', None))
    return train(model, data_engine_module, trainer.f_grad_shared, trainer.f_update, f_log_probs, f_alphas,
                 get_params, set_params, model_options, save_model_dir, use_noise=use_noise, **kw)
