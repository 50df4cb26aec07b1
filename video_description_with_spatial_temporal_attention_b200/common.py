"""Host utilities with the reference's names (reference ``common.py``): the
module-global numpy RNG the initialisers draw from, the weight initialisers,
parameter (un)zipping and minibatch indexing.  Host-side bookkeeping only."""
from __future__ import annotations

from collections import OrderedDict

import numpy


def get_two_rngs(seed=None):
    """reference common.py:16-23.  The second generator (Theano MRG stream in the
    reference) is a torch Philox seed here: dropout masks are drawn on the device
    (decision D2) and are not bit-comparable with MRG31k3p."""
    seed = 1234 if seed is None else seed
    return numpy.random.RandomState(seed), seed


rng_numpy, rng_seed = get_two_rngs()


def ortho_weight(ndim):
    """Left singular vectors of a Gaussian square matrix (common.py:110-122)."""
    u, _, _ = numpy.linalg.svd(rng_numpy.randn(ndim, ndim))
    return u.astype('float32')


def norm_weight(nin, nout=None, scale=0.01, ortho=True):
    """0.01*N(0,1), or orthogonal when square and ortho (common.py:124-134)."""
    nout = nin if nout is None else nout
    if nout == nin and ortho:
        return ortho_weight(nin)
    return (scale * rng_numpy.randn(nin, nout)).astype('float32')


def zipp(params, tparams):
    """push host values into the shared parameters (common.py:78-81)"""
    for k, v in params.items():
        tparams[k].set_value(v)


def unzip(zipped):
    """pull the shared parameters back to host numpy (common.py:84-88)"""
    return OrderedDict((k, v.get_value()) for k, v in zipped.items())


def itemlist(tparams):
    return [v for _, v in tparams.items()]


def flatten_list_of_list(l):
    return [x for sub in l for x in sub]


def generate_minibatch_idx(dataset_size, minibatch_size):
    """[m1, ..., mk] lists of indices; a shorter last one when uneven
    (common.py:297-311)."""
    assert dataset_size >= minibatch_size
    idx = list(range(dataset_size))
    out = [idx[i:i + minibatch_size] for i in range(0, dataset_size - dataset_size % minibatch_size,
                                                     minibatch_size)]
    if dataset_size % minibatch_size:
        out.append(idx[-(dataset_size % minibatch_size):])
    return out
