"""Parameter update of the training step on the device (second half; the gradients come from
stat_grad_shared, driven by train.Trainer): the reference's `f_update` for adam / adadelta
(common.py:178-230), global-norm clipping (model_attention.py:1194-1203) and the data-parallel
gradient all-reduce, all over ONE flat fp32 buffer in init_params order.

    flat = FlatParams(params, device)            # parameters as views into one buffer
    opt  = Adam(flat)                            # or Adadelta(flat)
    ...                                          # gradients written into opt.grads (flat, same layout)
    allreduce_grads(opt.grads)                   # SUM over ranks (no-op without a process group)
    opt.clip(clip_c)                             # after the all-reduce, as SURVEY 8(e) requires
    opt.f_update(lr)                             # the reference signature; lr is ignored by adam there too
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy
import torch

from . import _lib
from ._lib import check


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatParams(object):
    """All parameters in one contiguous fp32 device buffer, init_params order (the order of the
    reference's gradient list, `itemlist(tparams)`); `views[name]` are tensors into it."""

    def __init__(self, params, device=None):
        if not torch.cuda.is_available():
            raise _lib.StatError('no CUDA device: the optimizer kernels have no CPU path')
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.shapes = OrderedDict((k, tuple(numpy.asarray(v).shape)) for k, v in params.items())
        self.offsets, n = OrderedDict(), 0
        for k, shp in self.shapes.items():
            self.offsets[k] = n
            n += int(numpy.prod(shp)) if len(shp) else 1
        self.n = n
        self.flat = torch.empty(n, dtype=torch.float32, device=self.device)
        self.views = OrderedDict()
        for k, v in params.items():
            o, cnt = self.offsets[k], int(numpy.prod(self.shapes[k])) if len(self.shapes[k]) else 1
            self.flat[o:o + cnt].copy_(torch.from_numpy(numpy.asarray(v, 'float32').reshape(-1).copy()))
            self.views[k] = self.flat[o:o + cnt].view(self.shapes[k] if len(self.shapes[k]) else ())

    def unzip(self):
        """host copies, as common.unzip returns them"""
        return OrderedDict((k, v.detach().cpu().numpy()) for k, v in self.views.items())

    def zeros_like(self):
        return torch.zeros_like(self.flat)


def allreduce_grads(grads, group=None):
    """SUM of the flat gradient buffer over the ranks of the process group (NCCL over NVLink on GPUs,
    gloo in the CPU tests): one collective per step.  With the mean-over-global-batch NLL each rank
    scales its local sum by 1 / B_global beforehand, the coverage regulariser is a plain sum over the
    batch, weight decay is added once.  No-op when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=group)
    return grads


class _Optimizer(object):
    def __init__(self, flat):
        self.lib = _lib.load()
        self.p = flat
        self.grads = flat.zeros_like()
        self._scratch = torch.empty((int(self.lib.stat_clip_scratch_bytes()) + 3) // 4, dtype=torch.float32,
                                    device=flat.device)
        self._g2 = torch.zeros(2, dtype=torch.float32, device=flat.device)

    def clip(self, clip_c):
        """In place on self.grads; returns (g2, factor) as a device tensor of two floats."""
        check(self.lib.stat_grad_clip(_ptr(self.grads), self.p.n, float(clip_c), _ptr(self._scratch), _ptr(self._g2),
                                      _stream()))
        return self._g2


class Adam(_Optimizer):
    """common.py:197-230"""

    def __init__(self, flat):
        super(Adam, self).__init__(flat)
        self.m, self.v = flat.zeros_like(), flat.zeros_like()
        self.i = 0

    def f_update(self, lr=None):
        self.i += 1
        check(self.lib.stat_adam_step(_ptr(self.p.flat), _ptr(self.grads), _ptr(self.m), _ptr(self.v), self.p.n, self.i,
                                      _stream()))
        return []


class Adadelta(_Optimizer):
    """common.py:178-195"""

    def __init__(self, flat):
        super(Adadelta, self).__init__(flat)
        self.rg2, self.ru2 = flat.zeros_like(), flat.zeros_like()

    def grad_shared(self):
        """the running-gradient update that rides on f_grad_shared in the reference"""
        check(self.lib.stat_adadelta_step(_ptr(self.p.flat), _ptr(self.grads), _ptr(self.rg2), _ptr(self.ru2), self.p.n,
                                          0, _stream()))

    def f_update(self, lr=None):
        check(self.lib.stat_adadelta_step(_ptr(self.p.flat), _ptr(self.grads), _ptr(self.rg2), _ptr(self.ru2), self.p.n,
                                          1, _stream()))
        return []
