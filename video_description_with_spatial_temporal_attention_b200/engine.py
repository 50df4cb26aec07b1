"""Device engine: owns the packed parameters and the per-batch workspace (torch
tensors as the container for device memory) and drives ``libstat_b200.so``.

This is plumbing between the reference-shaped host API (``model_attention.py``
in this package) and the C ABI; all arithmetic happens in the CUDA kernels.
"""
from __future__ import annotations

import atexit
import ctypes as C
import functools
import os
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import StatDims, StatFwdBlocks, StatParams, check


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# Engines alive in this process.  Their captured CUDA graphs are released (after a device
# synchronisation) before the interpreter starts tearing modules down: graph-exec objects destroyed in
# arbitrary order during finalisation, possibly after the CUDA context, are a known way to die at exit.
_ENGINES = weakref.WeakSet()


def _release_graphs():
    try:
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            torch.cuda.synchronize()
    except Exception:
        pass
    for e in list(_ENGINES):
        e.__dict__.get('_stream_slots', {}).clear()
        e._graphs.clear()


atexit.register(_release_graphs)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device(fn):
    """Run an Engine method with the engine's device current: the ABI enqueues on the CURRENT device's
    stream and the library picks its side streams / events by cudaGetDevice(), so an Engine created for
    cuda:1 must not launch while cuda:0 is current."""
    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **kw)
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapper


class Engine(object):
    """One parameter set on one GPU."""

    def __init__(self, options, device=None):
        if not torch.cuda.is_available():
            raise _lib.StatError('no CUDA device: the STAT decoder has no CPU path')
        self.lib = _lib.load()
        self.options = dict(options)
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        o = self.options
        if not o.get('global_proj') and o['ctxg_dim'] != o['dim']:
            raise ValueError('the reference graph needs ctxg_dim == dim; set global_proj=True otherwise')
        if not o.get('use_dropout', True):
            # the kernels implement the dropout graph (0.5 factors at eval, masks when training); the
            # reference's use_dropout=False branch calls _step with a stale signature and cannot run
            # (model_attention.py:480-488)
            raise ValueError('use_dropout=False is not supported (the reference graph itself raises on that branch)')
        self.flags = _lib.flags_of(o)
        # L2 residency of the projected context blocks across the decode steps: the attention kernel copies them with
        # evict_last priority and streams everything else evict_first.  A persisting-L2 carve-out
        # (cudaLimitPersistingL2CacheSize) does NOT help: bulk copies with an evict_last hint are not "persisting"
        # accesses, so the set-aside only shrinks the L2 everything competes for (measured in round 2: attention
        # 22.1 us per launch with the 79 MB maximum, 21.1 with 40 MB, 20.9 with none).  STAT_L2_PERSIST=<MB> sets one.
        mb = os.environ.get('STAT_L2_PERSIST')
        with torch.cuda.device(self.device):
            self.l2_persist_bytes = int(self.lib.stat_set_l2_persist(0 if mb is None else int(mb) << 20))
        self.prepared = None
        self._dev_params = None
        self._ws = {}
        self._pinned = {}
        self._graphs = {}
        _ENGINES.add(self)

    # ---- dims / buffers --------------------------------------------------
    def dims(self, B, T, R):
        o = self.options
        return StatDims(B=B, T=T, R=R, Dg=o['ctxg_dim'], Dm=o['ctxm_dim'], Dr=o['ctxl_dim'], H=o['dim'],
                        E=o['dim_word'], V=o['n_words'], flags=self.flags)

    @_on_device
    def workspace(self, B, T, R, rows=None, tag=None):
        """`tag`: callers that keep results cached in a workspace across calls (build_sampler's per-clip
        context) pass a private tag so that no other user of the same shape overwrites it."""
        rows = B if rows is None else rows
        key = (B, T, R, rows) if tag is None else (B, T, R, rows, tag)
        ws = self._ws.get(key)
        if ws is None:
            d = self.dims(B, T, R)
            # a workspace made for `rows` decode rows also serves any smaller row count
            n = max(self.lib.stat_workspace_bytes(C.byref(d), r)
                    for r in (range(1, rows + 1) if rows != B else (rows,)))
            if n == 0:
                check(-1)
            ws = torch.zeros((n + 3) // 4, dtype=torch.float32, device=self.device)
            self._ws[key] = ws
        return ws

    def region(self, ws, B, T, R, name, rows=None):
        d = self.dims(B, T, R)
        off, nb = C.c_size_t(), C.c_size_t()
        check(self.lib.stat_workspace_region(C.byref(d), B if rows is None else rows, name.encode(),
                                             C.byref(off), C.byref(nb)))
        return ws[off.value // 4:(off.value + nb.value) // 4]

    def to_device(self, a, dtype, name=None):
        """numpy / torch (host or device) -> contiguous device tensor.  Host arrays go
        through a cached pinned staging buffer so the copy is a real async H2D."""
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a)
        if t.dtype != dtype:
            t = t.to(dtype)
        key = (name, tuple(t.shape), dtype)
        if name is not None:
            pin = self._pinned.get(key)
            if pin is None:
                pin = torch.empty(t.shape, dtype=dtype).pin_memory()
                self._pinned[key] = pin
            pin.copy_(t)
            t = pin
        return t.to(self.device, non_blocking=True)

    # ---- parameters --------------------------------------------------------
    @_on_device
    def set_params(self, params):
        """params: mapping name -> numpy / torch array in the reference's layout
        (init_params, model_attention.py:518-581)."""
        o = self.options
        dev = {}
        for k, v in params.items():
            t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(v, 'float32')))
            dev[k] = t.to(device=self.device, dtype=torch.float32).contiguous().reshape(-1)
        sp = StatParams()
        for f in _lib.PARAM_FIELDS:
            if f in dev:
                setattr(sp, f, dev[f].data_ptr())
        d = self.dims(1, 1, 1)
        n = self.lib.stat_prepared_bytes(C.byref(d))
        if n == 0:
            check(-1)
        if self.prepared is None or self.prepared.numel() * 4 < n:
            self.prepared = torch.empty((n + 3) // 4, dtype=torch.float32, device=self.device)
        check(self.lib.stat_prepare_params(C.byref(d), C.byref(sp), _ptr(self.prepared), _stream()))
        self._dev_params = dev      # keep alive until the packing kernels have run
        self.__dict__.get('_stream_slots', {}).clear()      # (they hold references to the captured graphs)
        self._graphs.clear()

    def _need_params(self):
        if self.prepared is None:
            raise _lib.StatError('set_params() has not been called')

    # ---- K0 ----------------------------------------------------------------
    @_on_device
    def precompute(self, ctxg, mask_ctxg, ctxl, ctxm, rows=None, ws_tag=None):
        """ctxg (B,T,Dg) mask (B,T) ctxl (B,T,R,Dr) ctxm (B,T,Dm), device fp32 tensors."""
        self._need_params()
        B, T = ctxg.shape[0], ctxg.shape[1]
        R = ctxl.shape[2]
        ws = self.workspace(B, T, R, rows, tag=ws_tag)
        d = self.dims(B, T, R)
        check(self.lib.stat_precompute(C.byref(d), _ptr(self.prepared), _ptr(ctxg), _ptr(mask_ctxg), _ptr(ctxl),
                                       _ptr(ctxm), _ptr(ws), _stream()))
        return ws, d

    # ---- f_init ----------------------------------------------------------------
    @_on_device
    def init_state(self, ctxg, mask_ctxg):
        self._need_params()
        B, T = ctxg.shape[0], ctxg.shape[1]
        ws = self.workspace(B, T, 1)
        d = self.dims(B, T, 1)
        h0 = torch.empty((B, d.H), dtype=torch.float32, device=self.device)
        c0 = torch.empty_like(h0)
        check(self.lib.stat_init_state(C.byref(d), _ptr(self.prepared), _ptr(ctxg), _ptr(mask_ctxg), _ptr(ws),
                                       _ptr(h0), _ptr(c0), _stream()))
        return h0, c0

    # ---- f_log_probs ---------------------------------------------------------
    @_on_device
    def forward_teacher(self, ws, d, x, mask, dp_gates=None, dp_h=None, dp_z=None, want_alphas=False,
                        want_h=False):
        L, B = x.shape
        T, R, H = d.T, d.R, d.H
        out = torch.empty(B, dtype=torch.float32, device=self.device)
        al = ag = am = alt = hh = None
        if want_alphas:
            al = torch.empty((L, B, T, R), dtype=torch.float32, device=self.device)
            ag = torch.empty((L, B, T), dtype=torch.float32, device=self.device)
            am = torch.empty_like(ag)
            alt = torch.empty_like(ag)
        if want_h:
            hh = torch.empty((L, B, H), dtype=torch.float32, device=self.device)
        check(self.lib.stat_forward_teacher(C.byref(d), _ptr(self.prepared), _ptr(ws), L, _ptr(x), _ptr(mask),
                                            _ptr(dp_gates), _ptr(dp_h), _ptr(dp_z), _ptr(out), _ptr(al), _ptr(ag),
                                            _ptr(am), _ptr(alt), _ptr(hh), _stream()))
        return out, (al, ag, am, alt), hh

    # ---- f_grad_shared (gradients only; clipping / update live in optim.py) ----------
    @_on_device
    def grad_shared(self, ws, d, batch_dev, alphas, h_all, grad_views, inv_batch, alpha_c=0., decay_c=0.,
                    dp_gates=None, dp_h=None, dp_z=None):
        """Gradients of the training cost (model_attention.py:1129-1147, :1193) for the batch whose forward
        (precompute + forward_teacher with alphas and hidden states) has just run in `ws`.
        batch_dev = (x, mask, ctxg, mask_ctxg, ctxl, ctxm) device tensors; grad_views: name -> writable fp32
        device tensor per parameter (e.g. views into the optimizer's flat gradient buffer)."""
        self._need_params()
        x, mask, ctxg, mask_ctxg, ctxl, ctxm = batch_dev
        L, B = x.shape
        T, R = d.T, d.R
        blocks = StatFwdBlocks()
        for name in StatFwdBlocks.FIELDS:
            setattr(blocks, name, self.region(ws, B, T, R, 'h0' if name == 'h0c0' else name).data_ptr())
        sp, sg = StatParams(), StatParams()
        for f in _lib.PARAM_FIELDS:
            if f in self._dev_params:
                setattr(sp, f, self._dev_params[f].data_ptr())
                setattr(sg, f, grad_views[f].data_ptr())
        n = self.lib.stat_grad_workspace_bytes(C.byref(d), L)
        if n == 0:
            check(-1)
        key = ('grad', B, T, R, L)
        gws = self._ws.get(key)
        if gws is None:
            gws = torch.empty((n + 3) // 4, dtype=torch.float32, device=self.device)
            self._ws[key] = gws
        al, ag, am, alt = alphas
        check(self.lib.stat_grad_shared(C.byref(d), C.byref(sp), C.byref(blocks), L, _ptr(x), _ptr(mask), _ptr(ctxg),
                                        _ptr(mask_ctxg), _ptr(ctxl), _ptr(ctxm), _ptr(dp_gates), _ptr(dp_h),
                                        _ptr(dp_z), _ptr(al), _ptr(ag), _ptr(am), _ptr(alt), _ptr(h_all),
                                        float(inv_batch), float(alpha_c), float(decay_c), C.byref(sg), _ptr(gws),
                                        _stream()))

    # ---- greedy ----------------------------------------------------------------
    @_on_device
    def decode_greedy(self, ws, d, maxlen, out=None):
        B = d.B
        if out is None:
            out = (torch.empty((B, maxlen), dtype=torch.int64, device=self.device),
                   torch.empty(B, dtype=torch.int32, device=self.device),
                   torch.empty(B, dtype=torch.float32, device=self.device))
        tokens, lengths, scores = out
        check(self.lib.stat_decode_greedy(C.byref(d), _ptr(self.prepared), _ptr(ws), maxlen, _ptr(tokens),
                                          _ptr(lengths), _ptr(scores), _stream()))
        return tokens, lengths, scores

    # ---- beam search ---------------------------------------------------------
    @_on_device
    def decode_beam(self, ws, d, k, maxlen):
        """ws: precompute(..., rows=B*k).  -> tokens (B,k,maxlen) i64, lengths (B,k) i32,
        scores (B,k) f32, count (B,) i32, device tensors in the reference's hypothesis order."""
        B = d.B
        if self.lib.stat_workspace_bytes(C.byref(d), B * k) > ws.numel() * 4:
            raise _lib.StatError('workspace too small for %d x %d beam rows' % (B, k))
        tokens = torch.empty((B, k, maxlen), dtype=torch.int64, device=self.device)
        lengths = torch.empty((B, k), dtype=torch.int32, device=self.device)
        scores = torch.empty((B, k), dtype=torch.float32, device=self.device)
        count = torch.empty(B, dtype=torch.int32, device=self.device)
        check(self.lib.stat_decode_beam(C.byref(d), _ptr(self.prepared), _ptr(ws), k, maxlen, _ptr(tokens),
                                        _ptr(lengths), _ptr(scores), _ptr(count), _stream()))
        return tokens, lengths, scores, count

    @_on_device
    def beam_captions(self, ctxg, mask_ctxg, ctxl, ctxm, k, maxlen, use_graph=False):
        """Features on the device -> decode_beam outputs (K0 + maxlen beam steps).  With use_graph the
        launch sequence is captured once per shape and replayed (outputs are then reused buffers)."""
        self._need_params()
        B, T, R = ctxg.shape[0], ctxg.shape[1], ctxl.shape[2]
        if not use_graph:
            ws, d = self.precompute(ctxg, mask_ctxg, ctxl, ctxm, rows=B * k)
            return self.decode_beam(ws, d, k, maxlen)
        key = ('beam', B, T, R, k, maxlen)
        g = self._graphs.get(key)
        if g is None:
            st = [torch.empty_like(t) for t in (ctxg, mask_ctxg, ctxl, ctxm)]
            for dst, src in zip(st, (ctxg, mask_ctxg, ctxl, ctxm)):
                dst.copy_(src)
            holder = {}

            def run():
                ws, d = self.precompute(*st, rows=B * k)
                holder['out'] = self.decode_beam(ws, d, k, maxlen)
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                run()                       # warm-up: function attributes, workspace allocation
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run()
            g = (graph, st, holder['out'])
            self._graphs[key] = g
        graph, st, out = g
        for dst, src in zip(st, (ctxg, mask_ctxg, ctxl, ctxm)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        graph.replay()
        return out

    # ---- f_next ----------------------------------------------------------------
    @_on_device
    def step(self, ws, d, x, h, c, row_clip=None):
        rows = x.shape[0]
        if self.lib.stat_workspace_bytes(C.byref(d), rows) > ws.numel() * 4:
            raise _lib.StatError('workspace too small for %d decode rows' % rows)
        probs = torch.empty((rows, d.V), dtype=torch.float32, device=self.device)
        h2 = torch.empty((rows, d.H), dtype=torch.float32, device=self.device)
        c2 = torch.empty_like(h2)
        check(self.lib.stat_step(C.byref(d), _ptr(self.prepared), _ptr(ws), rows, _ptr(row_clip), _ptr(x), _ptr(h),
                                 _ptr(c), _ptr(probs), _ptr(h2), _ptr(c2), _stream()))
        return probs, h2, c2

    # ---- whole-batch greedy captioning, optionally replayed from a CUDA graph ------
    @_on_device
    def greedy_captions(self, ctxg, mask_ctxg, ctxl, ctxm, maxlen, use_graph=True, slot=0):
        """Features already on the device -> (tokens (B,maxlen) i64, lengths (B,) i32,
        scores (B,) f32), all device tensors.  Runs K0 + the maxlen-step decode; with
        use_graph the launch sequence is captured once per shape (and `slot`: independent copies of the
        graph with their own static input / output tensors, for pipelines that fill one set of inputs while
        another is being decoded) and replayed."""
        self._need_params()
        B, T = ctxg.shape[0], ctxg.shape[1]
        R = ctxl.shape[2]
        if not use_graph:
            ws, d = self.precompute(ctxg, mask_ctxg, ctxl, ctxm)
            return self.decode_greedy(ws, d, maxlen)
        key = ('greedy', B, T, R, maxlen) if slot == 0 else ('greedy', B, T, R, maxlen, slot)
        g = self._graphs.get(key)
        if g is None:
            st = dict(ctxg=torch.empty_like(ctxg), mask=torch.empty_like(mask_ctxg), ctxl=torch.empty_like(ctxl),
                      ctxm=torch.empty_like(ctxm))
            for k, v in (('ctxg', ctxg), ('mask', mask_ctxg), ('ctxl', ctxl), ('ctxm', ctxm)):
                st[k].copy_(v)
            out = (torch.empty((B, maxlen), dtype=torch.int64, device=self.device),
                   torch.empty(B, dtype=torch.int32, device=self.device),
                   torch.empty(B, dtype=torch.float32, device=self.device))

            def run():
                ws, d = self.precompute(st['ctxg'], st['mask'], st['ctxl'], st['ctxm'])
                self.decode_greedy(ws, d, maxlen, out)
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                run()                       # warm-up: function attributes, workspace allocation
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run()
            g = (graph, st, out)
            self._graphs[key] = g
        graph, st, out = g
        if ctxg.data_ptr() != st['ctxg'].data_ptr():
            st['ctxg'].copy_(ctxg, non_blocking=True)
            st['mask'].copy_(mask_ctxg, non_blocking=True)
            st['ctxl'].copy_(ctxl, non_blocking=True)
            st['ctxm'].copy_(ctxm, non_blocking=True)
        graph.replay()
        return out

    def caption_stream(self, host_batches, maxlen, depth=2):
        """Greedy captions for a stream of host-resident batches, pipelined: the H2D copy of
        batch i+1 (copy stream, pinned host memory -> the static inputs of graph copy (i+1) % depth)
        overlaps the decode of batch i (compute stream: replay of graph copy i % depth, D2H of the
        captions).  One captured graph per pipeline slot, each with its own input / output tensors:
        the copies land where the graph reads them, nothing is staged on the device (a device-to-device
        staging copy on the compute stream cost 8 % of the H2D rate: tools/e2e_probe.py).

        host_batches: iterable of (ctxg, mask_ctxg, ctxl, ctxm) float32 torch tensors on the
        host (pinned memory for a truly asynchronous copy), all of one shape.  Yields
        (tokens (B,maxlen) int64, lengths (B,) int32, scores (B,) float32) numpy arrays, in
        order, each as soon as its batch has finished."""
        self._need_params()
        if torch.cuda.current_device() != self.device.index:
            raise _lib.StatError('caption_stream: make %s the current device first (torch.cuda.set_device)' % self.device)
        compute = torch.cuda.current_stream()
        copy = self._copy_stream = getattr(self, '_copy_stream', None) or torch.cuda.Stream(device=self.device)
        slots = None
        pending = []                      # (slot, done_event)

        def collect(slot, ev):
            ev.synchronize()
            return tuple(t.numpy().copy() for t in slots[slot]['out'])

        for i, hb in enumerate(host_batches):
            ctxg, mask, ctxl, ctxm = hb
            if slots is None:
                # the pipeline slots (captured graphs, pinned result buffers, events) are built on the first
                # call for a shape and kept on the engine: later streams start copying at once
                B, T, R = ctxg.shape[0], ctxg.shape[1], ctxl.shape[2]
                skey = ('stream', B, T, R, maxlen, depth)
                cache = self.__dict__.setdefault('_stream_slots', {})
                slots = cache.get(skey)
                if slots is None:
                    dev0 = [t.to(self.device) for t in hb]
                    slots = []
                    for k in range(depth):
                        self.greedy_captions(*dev0, maxlen=maxlen, use_graph=True, slot=k)      # capture once per slot
                        graph, st, gout = self._graphs[('greedy', B, T, R, maxlen) if k == 0 else
                                                       ('greedy', B, T, R, maxlen, k)]
                        slots.append(dict(graph=graph, inputs=[st['ctxg'], st['mask'], st['ctxl'], st['ctxm']],
                                          gout=gout,
                                          out=[torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in gout],
                                          ready=torch.cuda.Event(), done=torch.cuda.Event()))
                    del dev0
                    cache[skey] = slots
                for sl in slots:
                    sl['done'].record(compute)
            slot = i % depth
            sl = slots[slot]
            if len(pending) >= depth:                       # the slot's previous results must be read first
                yield collect(*pending.pop(0))
            with torch.cuda.stream(copy):
                copy.wait_event(sl['done'])                 # the slot's previous decode has read its inputs
                for dst, src in zip(sl['inputs'], hb):
                    dst.copy_(src, non_blocking=True)
                sl['ready'].record(copy)
            compute.wait_event(sl['ready'])
            sl['graph'].replay()
            for dst, src in zip(sl['out'], sl['gout']):
                dst.copy_(src, non_blocking=True)
            sl['done'].record(compute)
            pending.append((slot, sl['done']))
        while pending:
            yield collect(*pending.pop(0))

    def graph_inputs(self, B, T, R, maxlen, slot=0):
        """The static input tensors of the captured greedy graph (fill these in place to
        avoid the device-to-device staging copy)."""
        g = self._graphs.get(('greedy', B, T, R, maxlen) if slot == 0 else ('greedy', B, T, R, maxlen, slot))
        return None if g is None else g[1]

    # ---- the attention fragment alone (timing) ------------------------------------
    @_on_device
    def attention(self, ws, d, rows=None, row_clip=None):
        check(self.lib.stat_attention(C.byref(d), _ptr(self.prepared), _ptr(ws), d.B if rows is None else rows,
                                      _ptr(row_clip), _stream()))

    # ---- in-situ phase timing ------------------------------------------------------
    @_on_device
    def profile(self, fn):
        """Run fn() with the library's per-phase CUDA-event timing on; returns
        {phase: (total_ms, launches_groups)}.  Not for use under graph capture."""
        check(self.lib.stat_profile_enable(1))
        try:
            fn()
            n = self.lib.stat_profile_phases()
            ms = (C.c_float * n)()
            cnt = (C.c_int * n)()
            check(self.lib.stat_profile_collect(ms, cnt, n))
        finally:
            self.lib.stat_profile_enable(0)
        return {self.lib.stat_profile_phase_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def launch_count(self):
        return int(self.lib.stat_launch_count())

    # ---- the dense primitive (tests) ---------------------------------------------
    @_on_device
    def gemm(self, A, Bt, bias=None, alpha=1.0, post=1.0, act=0, swap=False):
        M, K = A.shape
        N = Bt.shape[0]
        Cc = torch.empty((M, N), dtype=torch.float32, device=self.device)
        check(self.lib.stat_gemm(_ptr(A), A.stride(0), _ptr(Bt), Bt.stride(0), _ptr(Cc), N, M, N, K, _ptr(bias),
                                 alpha, post, act, 1 if swap else 0, _stream()))
        return Cc
