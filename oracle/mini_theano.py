"""A very small lazy-graph stand-in for the dozen Theano ops the reference's hot
path uses, evaluated with numpy.  TEST INFRASTRUCTURE ONLY (see
``oracle/ref_exec.py``): it exists so that the reference's *own source text*
(``/root/reference/model_attention.py``) can be executed in this container,
where Theano is not installable, to pin ``oracle/stat_oracle.py``.

Semantics follow Theano's as used by the reference:
* ``tensor.dot(a, b)`` contracts the last axis of ``a`` with the first of ``b``.
* ``nnet.softmax`` is the max-subtracted row softmax of a matrix.
* ``x[:, None]`` follows the *old* ``__getitem__`` (the pattern is built from
  the written indices only, so a trailing broadcastable axis that is not
  indexed is dropped).  The reference depends on this: in one-step mode its
  mask is ``alloc(1., n, 1)`` and ``m_[:, None] * c`` must stay ``(n, dim)``
  (model_attention.py:310-311, :454-457).
* ``scan`` calls the step function once per time-step on constant leaves.
Nothing here is differentiable; only forward evaluation is needed.
"""
from __future__ import annotations

import sys
import types

import numpy as np


class _Tag(object):
    pass


def _as_var(x):
    if isinstance(x, Var):
        return x
    return Const(x)


def _ev(x, env):
    if isinstance(x, Var):
        return x.ev(env)
    if isinstance(x, (list, tuple)):
        return type(x)(_ev(i, env) for i in x)
    if isinstance(x, slice):
        return slice(_ev(x.start, env), _ev(x.stop, env), _ev(x.step, env))
    return x


class Var(object):
    """Lazy tensor.  ``fn(env) -> ndarray``; ``ndim`` is known statically."""

    def __init__(self, fn, ndim, dtype=None, name=None):
        self._fn = fn
        self.ndim = ndim
        self.dtype = dtype
        self.name = name
        self.tag = _Tag()

    def ev(self, env):
        k = id(self)
        if k not in env:
            env[k] = self._fn(env)
        return env[k]

    # ---- arithmetic ------------------------------------------------------
    def _bin(self, other, op, rev=False):
        o = _as_var(other)
        a, b = (o, self) if rev else (self, o)

        def fn(env):
            x, y = a.ev(env), b.ev(env)
            r = op(x, y)
            # keep float32 graphs float32 when mixed with python scalars
            if isinstance(r, np.ndarray) and r.dtype == np.float64:
                fl = [v for v in (x, y) if isinstance(v, np.ndarray) and v.ndim > 0
                      and v.dtype.kind == 'f']
                if fl and all(v.dtype == np.float32 for v in fl):
                    r = r.astype(np.float32)
            return r
        return Var(fn, max(a.ndim, b.ndim))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.true_divide)
    def __rtruediv__(self, o): return self._bin(o, np.true_divide, True)
    __div__ = __truediv__
    def __pow__(self, o): return self._bin(o, np.power)
    def __rpow__(self, o): return self._bin(o, np.power, True)
    def __lt__(self, o): return self._bin(o, np.less)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __neg__(self): return Var(lambda env: -self.ev(env), self.ndim)

    # ---- shape -----------------------------------------------------------
    @property
    def shape(self):
        return ShapeVar(self)

    def reshape(self, shp, ndim=None):
        shp = list(shp)
        return Var(lambda env: self.ev(env).reshape([int(_ev(s, env)) for s in shp]),
                   len(shp))

    def flatten(self, ndim=1):
        assert ndim == 1
        return Var(lambda env: self.ev(env).reshape(-1), 1)

    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (list, tuple)):
            pattern = tuple(pattern[0])

        def fn(env):
            a = self.ev(env)
            kept = [p for p in pattern if p != 'x']
            for d in range(a.ndim):
                if d not in kept:
                    assert a.shape[d] == 1, 'dimshuffle can only drop size-1 axes'
            a = a.transpose(kept + [d for d in range(a.ndim) if d not in kept])
            a = a.reshape(a.shape[:len(kept)])
            idx = tuple(None if p == 'x' else slice(None) for p in pattern)
            return a[idx]
        return Var(fn, len(pattern))

    def sum(self, axis=None):
        nd = 0 if axis is None else self.ndim - 1
        return Var(lambda env: self.ev(env).sum(axis=axis), nd)

    def mean(self, axis=None):
        nd = 0 if axis is None else self.ndim - 1
        return Var(lambda env: self.ev(env).mean(axis=axis), nd)

    def argmax(self, axis=None):
        return Var(lambda env: self.ev(env).argmax(axis=axis), self.ndim - 1)

    def astype(self, dt):
        return Var(lambda env: self.ev(env).astype(dt), self.ndim)

    # ---- indexing --------------------------------------------------------
    def __getitem__(self, args):
        if not isinstance(args, tuple):
            args = (args,)
        nd = self.ndim
        covered = 0
        out_nd = 0
        adv = 0
        for a in args:
            if a is None:
                out_nd += 1
            elif isinstance(a, slice):
                covered += 1
                out_nd += 1
            elif isinstance(a, Var) and a.ndim > 0:
                covered += 1
                adv = max(adv, a.ndim)
            else:
                covered += 1                       # integer: axis removed
        has_new = any(a is None for a in args)
        # old-Theano newaxis semantics (see module docstring): when None is
        # used, axes that are not indexed are dropped by the dimshuffle and
        # must therefore be broadcastable (size 1).
        drop = has_new and covered < nd
        out_nd += adv + (0 if drop else nd - covered)
        base = self

        def fn(env):
            arr = base.ev(env)
            idx = tuple(_ev(a, env) for a in args)
            if drop:
                assert all(s == 1 for s in arr.shape[covered:]), \
                    'newaxis indexing would drop a non-broadcastable axis'
                arr = arr.reshape(arr.shape[:covered])
            return arr[idx]
        v = Var(fn, out_nd)
        v._sub = (base, args)
        return v


class Const(Var):
    def __init__(self, value):
        value = np.asarray(value)
        Var.__init__(self, lambda env: value, value.ndim, value.dtype)


class Shared(Var):
    def __init__(self, value, name=None):
        self._v = np.asarray(value)
        Var.__init__(self, lambda env: self._v, self._v.ndim, self._v.dtype, name)

    def get_value(self, borrow=False):
        return self._v

    def set_value(self, v):
        self._v = np.asarray(v, self._v.dtype)


class Placeholder(Var):
    def __init__(self, name, ndim, dtype):
        def fn(env):
            raise KeyError('input %r was not supplied' % name)
        Var.__init__(self, fn, ndim, dtype, name)


class ShapeVar(object):
    def __init__(self, v):
        self.v = v

    def __getitem__(self, i):
        return Var(lambda env: np.int64(self.v.ev(env).shape[i]), 0)

    def as_tuple(self, env):
        return self.v.ev(env).shape


def _shape_list(shape, env):
    out = []
    for s in shape:
        if isinstance(s, ShapeVar):
            out.extend(s.as_tuple(env))
        else:
            out.append(int(_ev(s, env)))
    return out


# ---------------------------------------------------------------------------
# theano.tensor
# ---------------------------------------------------------------------------

def _elem(f):
    def op(x):
        x = _as_var(x)
        return Var(lambda env: f(x.ev(env)), x.ndim)
    return op


def _mk_input(ndim):
    def ctor(name=None, dtype='float32'):
        return Placeholder(name, ndim, dtype)
    return ctor


def _dot(a, b):
    a, b = _as_var(a), _as_var(b)
    return Var(lambda env: np.dot(a.ev(env), b.ev(env)), a.ndim + b.ndim - 2)


def _alloc(val, *shape):
    def fn(env):
        shp = _shape_list(shape, env)
        v = _ev(val, env)
        dt = np.float32 if isinstance(v, float) else np.asarray(v).dtype
        return np.full(shp, v, dtype=dt)
    return Var(fn, len(shape))


def _zeros(shape, dtype='float32'):
    return Var(lambda env: np.zeros(_shape_list(shape, env), dtype), len(shape))


def _zeros_like(x):
    return Var(lambda env: np.zeros_like(x.ev(env)), x.ndim)


def _set_subtensor(sub, value):
    base, args = sub._sub
    value = _as_var(value)

    def fn(env):
        out = base.ev(env).copy()
        out[tuple(_ev(a, env) for a in args)] = value.ev(env)
        return out
    return Var(fn, base.ndim)


def _switch(c, a, b):
    c, a, b = _as_var(c), _as_var(a), _as_var(b)

    def fn(env):
        cv = c.ev(env)
        av, bv = a.ev(env), b.ev(env)
        return np.where(cv != 0, av, bv).astype(np.result_type(av, bv))
    return Var(fn, max(c.ndim, a.ndim, b.ndim))


def _softmax(x):
    assert x.ndim == 2, 'nnet.softmax takes a matrix'

    def fn(env):
        a = x.ev(env)
        e = np.exp(a - a.max(axis=1, keepdims=True))
        return e / e.sum(axis=1, keepdims=True)
    return Var(fn, 2)


def _sigmoid(x):
    x = _as_var(x)
    one = np.float32(1.)

    def fn(env):
        a = x.ev(env)
        o = a.dtype.type(1) if isinstance(a, np.ndarray) else one
        return o / (o + np.exp(-a))
    return Var(fn, x.ndim)


def _maximum(a, b):
    return _as_var(b)._bin(a, np.maximum, True) if not isinstance(a, Var) \
        else a._bin(b, np.maximum)


def _arange(n):
    return Var(lambda env: np.arange(int(_ev(n, env))), 1)


def _addbroadcast(x, *axes):
    return x


def _concatenate(lst, axis=0):
    lst = [_as_var(v) for v in lst]
    return Var(lambda env: np.concatenate([v.ev(env) for v in lst], axis=axis),
               lst[0].ndim)


class _RandomStreams(object):
    """Stands in for MRG_RandomStreams.  Only reached when use_noise != 0 (or
    for next_sample, which every caller of f_next discards), so the stream
    itself is not part of any parity claim."""

    def __init__(self, seed=1234):
        self.rng = np.random.RandomState(seed)

    def binomial(self, size=None, p=0.5, n=1, dtype='float32'):
        shp = size if isinstance(size, (tuple, list)) else (size,)
        return Var(lambda env: self.rng.binomial(n, p, _shape_list(shp, env)).astype(dtype),
                   size.v.ndim if isinstance(size, ShapeVar) else len(shp))

    def multinomial(self, pvals=None, **kw):
        def fn(env):
            p = pvals.ev(env).astype(np.float64)
            p = p / p.sum(axis=1, keepdims=True)
            return np.stack([self.rng.multinomial(1, r) for r in p])
        return Var(fn, 2)


# ---------------------------------------------------------------------------
# theano.scan / theano.function
# ---------------------------------------------------------------------------

def scan(fn, sequences=None, outputs_info=None, non_sequences=None, name=None,
         n_steps=None, profile=False, mode=None, strict=False, **kw):
    sequences = [_as_var(s) for s in (sequences or [])]
    non_sequences = [_as_var(s) for s in (non_sequences or [])]
    outputs_info = list(outputs_info or [])
    rec = [i for i, o in enumerate(outputs_info) if o is not None]
    n_out = len(outputs_info)
    cell = {}

    def run(env):
        key = ('scan', id(cell))
        if key in env:
            return env[key]
        seqs = [s.ev(env) for s in sequences]
        nons = [Const(s.ev(env)) for s in non_sequences]
        prev = [np.asarray(_as_var(outputs_info[i]).ev(env)) for i in rec]
        steps = int(_ev(n_steps, env)) if n_steps is not None else len(seqs[0])
        outs = [[] for _ in range(n_out)]
        for t in range(steps):
            args = [Const(s[t]) for s in seqs] + [Const(p) for p in prev] + nons
            r = fn(*args)
            if not isinstance(r, (list, tuple)):
                r = [r]
            sub = dict(env)          # outer leaves stay visible to closures
            vals = [np.asarray(_as_var(v).ev(sub)) for v in r]
            for i, v in enumerate(vals):
                outs[i].append(v)
            prev = [vals[i] for i in rec]
        env[key] = [np.stack(o) for o in outs]
        return env[key]

    # static ndim of each output: probe one symbolic call
    outs = []
    for i in range(n_out):
        def mk(i):
            return lambda env: run(env)[i]
        oi = outputs_info[i]
        nd = (_as_var(oi).ndim + 1) if oi is not None else 0
        outs.append(Var(mk(i), nd))
    return outs, {}


def function(inputs, outputs, name=None, on_unused_input=None, profile=False,
             mode=None, updates=None, **kw):
    single = not isinstance(outputs, (list, tuple))
    outs = [outputs] if single else list(outputs)

    def call(*args):
        assert len(args) == len(inputs), \
            '%s: expected %d inputs, got %d' % (name, len(inputs), len(args))
        env = {}
        for ph, a in zip(inputs, args):
            a = np.asarray(a)
            if ph.dtype is not None:
                a = a.astype(ph.dtype)
            assert a.ndim == ph.ndim, (ph.name, a.ndim, ph.ndim)
            env[id(ph)] = a
        res = [np.asarray(o.ev(env)) for o in outs]
        for sv, nv in (updates or []):
            sv.set_value(_as_var(nv).ev(env))
        return res[0] if single else res
    return call


def shared(value, name=None, **kw):
    return Shared(value, name)


def install():
    """Register fake ``theano`` modules in sys.modules; returns the root."""
    th = types.ModuleType('theano')
    tt = types.ModuleType('theano.tensor')
    nnet = types.ModuleType('theano.tensor.nnet')
    sandbox = types.ModuleType('theano.sandbox')
    rng_mrg = types.ModuleType('theano.sandbox.rng_mrg')

    tt.matrix = _mk_input(2)
    tt.vector = _mk_input(1)
    tt.tensor3 = _mk_input(3)
    tt.tensor4 = _mk_input(4)
    tt.dot = _dot
    tt.alloc = _alloc
    tt.zeros = _zeros
    tt.zeros_like = _zeros_like
    tt.set_subtensor = _set_subtensor
    tt.switch = _switch
    tt.tanh = _elem(np.tanh)
    tt.log = _elem(np.log)
    tt.exp = _elem(np.exp)
    tt.sqrt = _elem(np.sqrt)
    tt.sqr = _elem(np.square)
    tt.maximum = _maximum
    tt.arange = _arange
    tt.addbroadcast = _addbroadcast
    tt.concatenate = _concatenate
    tt._shared = shared
    nnet.softmax = _softmax
    nnet.sigmoid = _sigmoid
    nnet.relu = lambda x: _maximum(0., x)
    tt.nnet = nnet
    rng_mrg.MRG_RandomStreams = _RandomStreams
    sandbox.rng_mrg = rng_mrg
    th.tensor = tt
    th.sandbox = sandbox
    th.scan = scan
    th.function = function
    th.shared = shared
    th.config = types.SimpleNamespace(floatX='float32')
    for n, m in (('theano', th), ('theano.tensor', tt), ('theano.tensor.nnet', nnet),
                 ('theano.sandbox', sandbox), ('theano.sandbox.rng_mrg', rng_mrg)):
        sys.modules[n] = m
    return th


def uninstall():
    for n in ('theano', 'theano.tensor', 'theano.tensor.nnet', 'theano.sandbox',
              'theano.sandbox.rng_mrg'):
        sys.modules.pop(n, None)
