"""Run the reference's own hot-path source on the numpy stand-in for Theano.

TEST INFRASTRUCTURE ONLY; works only where ``/root/reference`` is mounted
(the authoring container), never on the GPU box.  It exists to PIN the oracle:
``tests/golden/make_golden.py`` uses it to write small input/output fixtures
which ``tests/test_oracle_golden.py`` replays against ``oracle/stat_oracle.py``
everywhere.

How: the text of ``/root/reference/common.py`` and of
``/root/reference/model_attention.py`` (lines before ``def train``: the class
header, parameter initialisers, ``lstm_cond_layer``, ``build_model``,
``build_sampler``, ``gen_sample``, ``pred_probs``) is read at run time, given
the minimal Python-2 -> 3 token fixes listed in ``_py3`` (print statement,
xrange, iteritems, cPickle, integer ``/`` in gen_sample, tab expansion), and
exec'd with ``oracle/mini_theano`` registered as ``theano``.  For options with
``global_proj`` (decision D1) the three commented-out ``ff_global`` statements of the
reference are un-commented in memory first (``_enable_ff_global``), so the D1 fixtures
come from the reference's own lines too.  No reference source is copied into this
repository.
"""
from __future__ import annotations

import os
import re
import sys
import types

import numpy as np

from . import mini_theano

REF = os.environ.get('STAT_REFERENCE_DIR', '/root/reference')


def mounted():
    """The reference sources are present (authoring container only)."""
    return os.path.isfile(os.path.join(REF, 'model_attention.py'))


def available():
    """Executing the reference's source text in-process is OPT-IN: it runs third-party code, so the
    build entry point and the default test run never do it.  Set STAT_RUN_REFERENCE=1 to enable the
    live re-execution tests (the fixture generators under tests/golden/ call _load() explicitly)."""
    return mounted() and os.environ.get('STAT_RUN_REFERENCE', '0') == '1'


def _py3(src):
    src = src.expandtabs(8)
    out = []
    for line in src.split('\n'):
        m = re.match(r'^(\s*)print\s*$', line)
        if m:
            line = m.group(1) + 'print()'
        else:
            m = re.match(r'^(\s*)print\s+(?!\()(.*?)(,?)\s*$', line)
            if m and not m.group(2).rstrip().endswith('\\'):
                end = ", end=' '" if m.group(3) else ''
                line = '%sprint(%s%s)' % (m.group(1), m.group(2), end)
        out.append(line)
    src = '\n'.join(out)
    src = src.replace('xrange(', 'range(')
    src = src.replace('.iteritems()', '.items()')
    src = src.replace('import cPickle as pkl', 'import pickle as pkl')
    src = src.replace('import cPickle, os', 'import pickle as cPickle, os')
    src = src.replace('from sklearn.cross_validation import KFold', '')
    # python-2 integer division of argsort ranks (model_attention.py:926)
    src = src.replace('trans_indices = ranks_flat / voc_size',
                      'trans_indices = ranks_flat // voc_size')
    return src


def _enable_ff_global(lines):
    """Decision D1 from the reference's own text: the `ff_global` layer (a tanh projection of the global
    features to `dim`) exists in /root/reference/model_attention.py only as three commented-out
    two-line statements -- the parameter (:553-554), its use in build_model (:661-662) and in
    build_sampler (:780-781).  Un-comment exactly those six lines, in memory."""
    out = list(lines)
    hits = 0
    for i, line in enumerate(out):
        if "prefix='ff_global'" in line and line.lstrip().startswith('#'):
            assert out[i - 1].lstrip().startswith('#') and "get_layer('ff')" in out[i - 1], out[i - 1]
            out[i - 1] = out[i - 1].replace('# ', '', 1)
            out[i] = out[i].replace('# ', '', 1)
            hits += 1
    assert hits == 3, 'expected the three commented ff_global sites, found %d' % hits
    return out


def _load(global_proj=False):
    th = mini_theano.install()
    for stub in ('data_engine', 'metrics'):
        if stub not in sys.modules:
            sys.modules[stub] = types.ModuleType(stub)
    with open(os.path.join(REF, 'common.py')) as f:
        csrc = _py3(f.read())
    common = types.ModuleType('common')
    common.__file__ = os.path.join(REF, 'common.py')
    exec(compile(csrc, common.__file__, 'exec'), common.__dict__)
    saved_common = sys.modules.get('common')
    sys.modules['common'] = common
    with open(os.path.join(REF, 'model_attention.py')) as f:
        lines = f.read().split('\n')
    cut = next(i for i, l in enumerate(lines) if re.match(r'\s+def train\(self', l))
    if global_proj:
        lines = _enable_ff_global(lines[:cut]) + lines[cut:]
    msrc = _py3('\n'.join(lines[:cut]))
    mod = types.ModuleType('ref_model_attention')
    mod.__file__ = os.path.join(REF, 'model_attention.py')
    try:
        exec(compile(msrc, mod.__file__, 'exec'), mod.__dict__)
    finally:
        if saved_common is None:
            sys.modules.pop('common', None)
        else:
            sys.modules['common'] = saved_common
    return th, common, mod


class RefModel(object):
    """The reference's Attention object with its compiled callables."""

    def __init__(self, options, params=None, quiet=True):
        import contextlib
        import io
        self.th, self.common, self.mod = _load(global_proj=bool(options.get('global_proj')))
        self.options = dict(options)
        self.options.setdefault('encoder', 'none')
        sink = io.StringIO() if quiet else sys.stdout
        with contextlib.redirect_stdout(sink):
            m = self.mod.Attention()
            for a in ('x_tv', 'mask_tv', 'ctxg_tv', 'ctxg_mask_tv', 'ctxl_tv',
                      'ctxl_mask_tv', 'ctxm_tv', 'ctxm_mask_tv'):
                setattr(m, a, None)
            self.common.rng_numpy.seed(1234)          # common.py:25
            self.init = m.init_params(self.options)
            self.params = self.init if params is None else params
            tparams = m.init_tparams(self.params)
            r = m.build_model(tparams, self.options)
            (trng, use_noise, x, mask, ctxg, mask_ctxg, ctxl, mask_ctxl, ctxm,
             mask_ctxm, alphals, alphags, alphams, alphalts, cost, extra) = r
            inps = [x, mask, ctxg, mask_ctxg, ctxl, mask_ctxl, ctxm, mask_ctxm]
            self.use_noise = use_noise
            # model_attention.py:1126
            self.f_log_probs = self.th.function(inps, -cost, name='f_log_probs')
            self.f_extra = self.th.function(
                inps, [extra[0], alphals, alphags, alphams, alphalts], name='f_extra')
            self.f_init, self.f_next = m.build_sampler(tparams, self.options,
                                                       use_noise, trng)
        self.model = m
        self.tparams = tparams

    def gen_sample(self, ctxg, ctxg_mask, ctxl, ctxl_mask, ctxm, ctxm_mask, k, maxlen):
        return self.model.gen_sample(None, self.f_init, self.f_next, ctxg, ctxg_mask,
                                     ctxl, ctxl_mask, ctxm, ctxm_mask, self.options,
                                     None, k, maxlen, False)
