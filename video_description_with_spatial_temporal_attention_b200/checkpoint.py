"""Checkpoint files of the reference, read and written with its own layout:

  model_current.npz / model_best_so_far.npz / model_best.npz
      numpy.savez(path, history_errs=..., [train_err=, valid_err=, test_err=,] **params)
      (model_attention.py:1388-1390, 1477-1490, 1545-1548)
  model_options.pkl
      the options dict, pickled (model_attention.py:1083-1084, 1491-1492); files written by the
      reference are Python-2 pickles and are read with latin-1 decoding.

``load_params`` applies the checks of the reference's loaders (metrics.py:30-38,
model_attention.py:1109-1113): every key of init_params must be in the archive with the same
shape; 0-d entries come back as ndarray.
"""
from __future__ import annotations

import os
import pickle
from collections import OrderedDict

import numpy

from . import common

EXTRA_KEYS = ('history_errs', 'train_err', 'valid_err', 'test_err', 'zipped_params')


def save_options(save_model_dir, model_options):
    with open(os.path.join(save_model_dir, 'model_options.pkl'), 'wb') as f:
        pickle.dump(dict(model_options), f, protocol=2)       # protocol 2: also readable by the reference (py2)


def load_options(path):
    if os.path.isdir(path):
        path = os.path.join(path, 'model_options.pkl')
    with open(path, 'rb') as f:
        try:
            return pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            return pickle.load(f, encoding='latin1')


def save_params(path, params, history_errs=(), **errs):
    """params: OrderedDict of numpy arrays, or the shared-parameter dict (unzipped here)."""
    first = next(iter(params.values()))
    if hasattr(first, 'get_value'):
        params = common.unzip(params)
    numpy.savez(path, history_errs=numpy.asarray(history_errs), **dict(errs), **params)


def load_params(path, params):
    """Fill `params` (the dict init_params returns) from an archive; returns it."""
    pp = numpy.load(path, allow_pickle=True, encoding='latin1')
    for k in list(params.keys()):
        if k not in pp:
            raise Exception('%s is not in the archive' % k)
        v = numpy.asarray(pp[k], dtype='float32')
        if v.shape != numpy.asarray(params[k]).shape:
            raise ValueError('%s: archive shape %s, model shape %s' % (k, v.shape, numpy.asarray(params[k]).shape))
        params[k] = v
    return params


def archive_extras(path):
    """The non-parameter entries of an archive (error history, final errors)."""
    pp = numpy.load(path, allow_pickle=True, encoding='latin1')
    return OrderedDict((k, pp[k]) for k in EXTRA_KEYS if k in pp)
