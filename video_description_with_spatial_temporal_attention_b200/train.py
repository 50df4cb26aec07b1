"""The training step of train() on the device: what the reference builds at
model_attention.py:1129-1147 (cost), :1193-1203 (gradients, global-norm clipping) and :1206-1209
(`f_grad_shared, f_update = optimizer(lr, tparams, grads, inps, cost, extra)`, common.py:178-230), and
calls at :1259-1278 as

    rvals = f_grad_shared(x, mask, ctxg, mask_ctxg, ctxl, mask_ctxl, ctxm, mask_ctxm)
    cost, alphas = rvals[0], rvals[2:6]
    f_update(lrate)

Parameters, gradients and optimizer state live in flat fp32 device buffers (optim.py); one step is
stat_prepare_params -> stat_precompute -> stat_forward_teacher -> stat_grad_shared -> (SUM all-reduce over
the data-parallel ranks) -> stat_grad_clip -> stat_adam_step / stat_adadelta_step.  Data parallelism
(SURVEY 8e): every rank scales its NLL sum by 1 / B_global, the coverage regulariser is a sum over clips,
weight decay is added on rank 0 only, clipping runs after the all-reduce, every rank applies the same update.
There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch

from . import optim
from ._lib import check
from .engine import Engine

__all__ = ['Trainer']


class Trainer(object):
    def __init__(self, params, options, optimizer='adam', alpha_c=0., decay_c=0., clip_c=0., device=None,
                 group=None, use_noise=True, seed=1234, sync_cost=True):
        self.options = dict(options)
        self.engine = Engine(self.options, device)
        self.flat = optim.FlatParams(params, self.engine.device)
        if optimizer == 'adam':
            self.opt = optim.Adam(self.flat)
        elif optimizer == 'adadelta':
            self.opt = optim.Adadelta(self.flat)
        else:
            raise ValueError('optimizer must be adam or adadelta (the two the reference config offers)')
        self.alpha_c, self.decay_c, self.clip_c = float(alpha_c), float(decay_c), float(clip_c)
        self.group = group
        self.use_noise = bool(use_noise)
        # data-parallel ranks must not draw identical dropout masks for their shards: seed + rank
        self.gen = torch.Generator(device=self.engine.device)
        self.gen.manual_seed(seed + self._world()[0])
        # sync_cost=False: f_grad_shared returns the cost as a 0-d device tensor instead of a python float, which
        # removes the one host synchronisation of the step (the reference's numpy return forces it, :1259-1262)
        self.sync_cost = bool(sync_cost)
        # measurement hook: CUDA events around the gradient all-reduce of every step (bench.py train_dp)
        self.time_allreduce = False
        self.allreduce_events = []
        self._dirty = True
        self.grad_views = OrderedDict()
        for k, shp in self.flat.shapes.items():
            o = self.flat.offsets[k]
            self.grad_views[k] = self.opt.grads[o:o + self.flat.views[k].numel()]
        self._cov = torch.zeros(4, dtype=torch.float32, device=self.engine.device)
        self.last = {}

    # ---- distributed helpers -------------------------------------------------------------------
    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def dropout_masks(self, L, B):
        """Bernoulli(0.5) keep masks (common.py:94-99, model_attention.py:469-477, :684-696) drawn on the device."""
        if not self.use_noise:
            return None, None, None
        H, E = self.options['dim'], self.options['dim_word']
        dev = self.engine.device
        # one Bernoulli draw straight into the float buffer of the three masks (contiguous chunks): 1 launch, one write
        n = L * B
        u = torch.empty(n * (4 * H + E), dtype=torch.float32, device=dev).bernoulli_(0.5, generator=self.gen)
        return (u[:n * 3 * H].view(L, B, 3 * H), u[n * 3 * H:n * 4 * H].view(L, B, H),
                u[n * 4 * H:].view(L, B, E))

    # ---- f_grad_shared -------------------------------------------------------------------------
    def f_grad_shared(self, x, mask, ctxg, mask_ctxg, ctxl, mask_ctxl, ctxm, mask_ctxm, dropout=None,
                      global_batch=None, return_grads=False):
        """Returns [cost, None, alphals, alphags, alphams, alphalts] (the reference's `probs` output, a
        (L*B, V) matrix nothing downstream reads, is not materialised); cost is a python float, the alphas
        device tensors.  The clipped, all-reduced gradients stay in the optimizer's flat buffer, as the
        reference stashes them in shared variables (common.py:198-201).  `dropout` = explicit
        (dp_gates, dp_h, dp_z) factors instead of drawn masks (parity tests).  return_grads=True appends host copies
        of the clipped gradients in init_params order, i.e. the reference's `rvals[6:]` (one device-to-host copy of
        every parameter-sized array per step: what train() pays for its NaN report, :1263-1269)."""
        eng = self.engine
        if self._dirty:
            eng.set_params(self.flat.views)
            self._dirty = False
        f32, i64 = torch.float32, torch.int64
        xd = eng.to_device(x, i64, 'x')
        md = eng.to_device(mask, f32, 'mask')
        gd = eng.to_device(ctxg, f32, 'ctxg')
        gmd = eng.to_device(mask_ctxg, f32, 'mask_ctxg')
        ld = eng.to_device(ctxl, f32, 'ctxl')
        mmd = eng.to_device(ctxm, f32, 'ctxm')
        L, B = xd.shape
        rank, world = self._world()
        inv_batch = 1.0 / float(global_batch if global_batch is not None else B * world)
        ws, d = eng.precompute(gd, gmd, ld, mmd)
        if dropout is not None:
            dpg, dph, dpz = (None if a is None else eng.to_device(a, f32) for a in dropout)
        else:
            dpg, dph, dpz = self.dropout_masks(L, B)
        lp, alphas, h_all = eng.forward_teacher(ws, d, xd, md, dpg, dph, dpz, want_alphas=True, want_h=True)
        eng.grad_shared(ws, d, (xd, md, gd, gmd, ld, mmd), alphas, h_all, self.grad_views, inv_batch,
                        alpha_c=self.alpha_c, decay_c=self.decay_c if rank == 0 else 0.,
                        dp_gates=dpg, dp_h=dph, dp_z=dpz)
        if self.time_allreduce:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            optim.allreduce_grads(self.opt.grads, self.group)
            e1.record()
            self.allreduce_events.append((e0, e1))
        else:
            optim.allreduce_grads(self.opt.grads, self.group)
        g2 = self.opt.clip(self.clip_c) if self.clip_c > 0. else None
        if isinstance(self.opt, optim.Adadelta):
            self.opt.grad_shared()
        # the cost itself (:1129-1147)
        cost = (-lp.double().sum()) * inv_batch
        if self.alpha_c > 0.:
            lib = eng.lib
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for i, a in enumerate(alphas):
                n = int(a.numel() // (a.shape[0] * a.shape[1]))
                check(lib.stat_alpha_coverage(C.c_void_p(a.data_ptr()), a.shape[0], a.shape[1], n,
                                              C.c_void_p(self.opt._scratch.data_ptr()),
                                              C.c_void_p(self._cov[i:].data_ptr()), stream))
            cost = cost + self.alpha_c * self._cov.double().sum()
        if world > 1:
            import torch.distributed as dist
            cost = cost.reshape(1)
            dist.all_reduce(cost, op=dist.ReduceOp.SUM, group=self.group)
            cost = cost[0]
        if self.decay_c > 0.:
            # sum of squares accumulated in fp64 in ONE pass over the parameters (no fp64 copy of the 52 M floats)
            cost = cost + self.decay_c * torch.linalg.vector_norm(self.flat.flat, 2, dtype=torch.float64).square()
        self.last = dict(logp=lp, g2=g2, tokens=md.sum())
        out = [float(cost) if self.sync_cost else cost, None] + list(alphas)
        if return_grads:
            out += list(self.grads().values())
        return out

    # ---- f_update ----------------------------------------------------------------------------
    def f_update(self, lr=None):
        """common.py:190-195 / :216-230: applies the stashed gradients; `lr` is accepted and (for adam, as in the
        reference) ignored."""
        self.opt.f_update(lr)
        self._dirty = True
        return []

    def grad_norm2(self):
        """Squared global norm of the last all-reduced gradient, before clipping (the first stage of the clip,
        model_attention.py:1194-1197) -- what train() checks for NaN / inf (the reference's grad_nan_report,
        :1263-1269).  None when clipping is off."""
        g2 = self.last.get('g2')
        return None if g2 is None else float(g2[0])

    def grads(self):
        """host copies of the current (clipped, all-reduced) gradients, init_params order"""
        return OrderedDict((k, v.detach().cpu().numpy().reshape(self.flat.shapes[k]))
                           for k, v in self.grad_views.items())

    def unzip(self):
        return self.flat.unzip()

    def load_params(self, params):
        """Overwrite the device parameters (e.g. `zipp(best_p, ...)`, model_attention.py:1522-1523); optimizer
        state is kept, like the reference's shared m / v / running averages."""
        import numpy
        for k, v in params.items():
            self.flat.views[k].copy_(torch.from_numpy(numpy.ascontiguousarray(numpy.asarray(v, 'float32')))
                                     .reshape(self.flat.views[k].shape))
        self._dirty = True

    # ---- hand-over to the reference-shaped validation / sampling callables -------------------------
    @classmethod
    def from_tparams(cls, tparams, options, **kw):
        """Trainer over the current values of the shared parameters (`unzip(tparams)`, common.py:84-88)."""
        from . import common
        return cls(common.unzip(tparams), options, **kw)

    def sync_tparams(self, tparams):
        """`zipp(self.unzip(), tparams)` (common.py:78-81): what train() gets for free from Theano's in-place updates
        of the shared variables -- call before pred_probs / gen_sample / numpy.savez(**unzip(tparams)) so that
        f_log_probs, f_init / f_next and the checkpoint see the trained values (one device-to-host copy of the
        parameters, at validation frequency)."""
        from . import common
        common.zipp(self.unzip(), tparams)
