"""Host-side mirror of the reference's callable surface for the caption-decoding
hot path (reference ``model_attention.py``): ``Attention.init_params`` /
``init_tparams`` / ``build_model`` / ``build_sampler`` / ``gen_sample`` /
``pred_probs`` and ``function(inps, -cost)`` for ``f_log_probs``, with the same
names, positional signatures and numpy in / numpy out conventions, so the
reference's ``train()`` / ``metrics.py`` call sites work unchanged (SURVEY §8b,
INTEGRATION.md).  All arithmetic runs in the sm_100a kernels behind
``libstat_b200.so``; nothing here computes on the CPU except the beam
bookkeeping that the reference also does on the host.
"""
from __future__ import annotations

import sys
from collections import OrderedDict

import numpy
import torch

from . import common
from .engine import Engine

BATCH_NAMES = ('x', 'mask', 'ctxg', 'mask_ctxg', 'ctxl', 'mask_ctxl', 'ctxm', 'mask_ctxm')


class Shared(object):
    """Stand-in for a Theano shared variable: get_value / set_value."""

    def __init__(self, value, name=None, owner=None):
        self.name = name
        self._owner = owner
        self._value = numpy.asarray(value)

    def get_value(self, borrow=False):
        return self._value

    def set_value(self, value, borrow=False):
        self._value = numpy.asarray(value, dtype=self._value.dtype)
        if self._owner is not None:
            self._owner.version += 1


class TParams(OrderedDict):
    """tparams: name -> Shared; bumps ``version`` whenever a value changes so the
    engine re-packs the device copy lazily."""

    def __init__(self):
        OrderedDict.__init__(self)
        self.version = 0
        self.engine = None
        self.engine_version = -1


class Handle(object):
    """Opaque stand-in for a symbolic variable returned by build_model."""

    def __init__(self, name, model=None, neg=False):
        self.name = name
        self.model = model
        self.neg = neg

    def __neg__(self):
        return Handle(self.name, self.model, not self.neg)

    def __repr__(self):
        return '<stat %s%s>' % ('-' if self.neg else '', self.name)


class _Model(object):
    """What build_model ties together: parameters, options, the noise switch."""

    def __init__(self, tparams, options, use_noise, seed):
        self.tparams = tparams
        self.options = options
        self.use_noise = use_noise
        self.gen = None
        self.seed = seed

    def engine(self):
        tp = self.tparams
        if tp.engine is None:
            tp.engine = Engine(self.options)
        if tp.engine_version != tp.version:
            tp.engine.set_params(OrderedDict((k, v.get_value()) for k, v in tp.items()))
            tp.engine_version = tp.version
        return tp.engine

    def dropout_masks(self, L, B, eng):
        """use_noise=1: Bernoulli(0.5) keep masks drawn on the device (D2).
        use_noise=0: None -> the kernels use the constant 0.5 (common.py:94-99)."""
        if not float(self.use_noise.get_value()):
            return None, None, None
        if self.gen is None:
            self.gen = torch.Generator(device=eng.device)
            self.gen.manual_seed(self.seed)
        H, E = self.options['dim'], self.options['dim_word']
        draw = lambda *s: (torch.rand(*s, device=eng.device, generator=self.gen) < 0.5).float()
        return draw(L, B, 3 * H), draw(L, B, H), draw(L, B, E)

    def run_teacher(self, batch, want_alphas=False):
        x, mask, ctxg, mask_ctxg, ctxl, mask_ctxl, ctxm, mask_ctxm = batch
        eng = self.engine()
        f32, i64 = torch.float32, torch.int64
        xd = eng.to_device(x, i64, 'x')
        md = eng.to_device(mask, f32, 'mask')
        gd = eng.to_device(ctxg, f32, 'ctxg')
        gmd = eng.to_device(mask_ctxg, f32, 'mask_ctxg')
        ld = eng.to_device(ctxl, f32, 'ctxl')
        mmd = eng.to_device(ctxm, f32, 'ctxm')
        # mask_ctxl / mask_ctxm are accepted and ignored like the reference
        ws, d = eng.precompute(gd, gmd, ld, mmd)
        dpg, dph, dpz = self.dropout_masks(xd.shape[0], xd.shape[1], eng)
        return eng.forward_teacher(ws, d, xd, md, dpg, dph, dpz, want_alphas=want_alphas)


def function(inps, out, **kw):
    """``theano.function(inps, -cost)`` of model_attention.py:1126 -> f_log_probs.
    Also accepts the alpha handles (or a list of handles) returned by build_model."""
    outs = out if isinstance(out, (list, tuple)) else [out]
    model = outs[0].model
    names = [o.name for o in outs]
    for o in outs:
        if o.name not in ('cost', 'alphals', 'alphags', 'alphams', 'alphalts'):
            raise NotImplementedError('function(): unsupported output %r' % o)

    def fn(*batch):
        if len(batch) != 8:
            raise TypeError('expected the 8 arrays of data_engine.prepare_data, got %d' % len(batch))
        want = any(n != 'cost' for n in names)
        lp, alphas, _ = model.run_teacher(batch, want_alphas=want)
        res = []
        for o in outs:
            if o.name == 'cost':
                v = lp.cpu().numpy()
                res.append(v if o.neg else -v)           # cost = -f_log_probs
            else:
                a = alphas[('alphals', 'alphags', 'alphams', 'alphalts').index(o.name)]
                res.append(a.cpu().numpy())
        return res if isinstance(out, (list, tuple)) else res[0]
    return fn


def draw_tokens(p, rng):
    """One multinomial draw per row of the probability matrix p (k, V) -> (k,) int64
    (`trng.multinomial(pvals=next_probs).argmax(1)`, model_attention.py:841)."""
    p = numpy.asarray(p, 'float64')
    c = numpy.cumsum(p, axis=1)
    u = rng.random_sample(p.shape[0]) * c[:, -1]
    return numpy.minimum((c < u[:, None]).sum(1), p.shape[1] - 1).astype('int64')


def validate_options(options):
    if options.get('ctx2out') and not options.get('ctxglm_dim', options['dim']) == options['dim']:
        raise ValueError('ctxglm_dim must equal dim (fusion is a sum, model_attention.py:430)')
    if not options.get('use_dropout', True):
        raise ValueError('use_dropout=False is not supported: the kernels implement the dropout graph, and the '
                         "reference's own no-dropout branch calls _step with a stale signature "
                         '(model_attention.py:480-488)')
    if options['dim_word'] > options['dim']:
        # model_attention.py:38-39
        print('WARNING: dim_word should only be as large as dim.')
    return options


class Attention(object):
    def __init__(self, channel=None):
        self.channel = channel
        self.engine = None          # the data engine, set by train() in the reference (:1087)

    # ---- parameters (model_attention.py:80-87, :180-282, :518-581) ---------------
    def param_init_fflayer(self, options, params, prefix='ff', nin=None, nout=None):
        params[prefix + '_W'] = common.norm_weight(nin, nout, scale=0.01)
        params[prefix + '_b'] = numpy.zeros((nout,), 'float32')
        return params

    def param_init_lstm_cond(self, options, params, nin, dim, dimctx=None, prefix='decoder'):
        nw = common.norm_weight
        z = lambda *s: numpy.zeros(s, 'float32')
        pre = prefix + '_'
        params[pre + 'W'] = numpy.concatenate([nw(nin, dim) for _ in range(4)], axis=1)
        params[pre + 'U'] = numpy.concatenate([common.ortho_weight(dim) for _ in range(4)], axis=1)
        params[pre + 'b'] = z(4 * dim)
        params[pre + 'Wc'] = nw(dim, 4 * dim)
        for n in ('Wcg_att', 'Wcm_att', 'Wclt_att'):
            params[pre + n] = nw(dim, ortho=False)
        for n in ('Wdg_att', 'Wdm_att', 'Wdlt_att'):
            params[pre + n] = nw(dim, dim)
        for n in ('bg_att', 'bm_att', 'blt_att'):
            params[pre + n] = z(dim)
        params[pre + 'Wcl_att'] = nw(dim, ortho=False)
        params[pre + 'Wdl_att'] = nw(dim, dim)
        params[pre + 'bl_att'] = z(dim)
        for n in ('g', 'm', 'lt', 'l'):
            params[pre + 'U%s_att' % n] = nw(dim, 1)
            params[pre + 'c%s_att' % n] = z(1)
        if options['selector']:
            params[pre + 'W_sel'] = nw(dim, 1)
            params[pre + 'b_sel'] = numpy.float32(0.)
        return params

    def init_params(self, options):
        o = options
        if not o.get('global_proj') and o['ctxg_dim'] != o['dim']:
            raise ValueError('the reference graph needs ctxg_dim == dim (SURVEY F3); '
                             'set global_proj=True to enable the ff_global layer')
        if o.get('encoder', 'none') != 'none' or o.get('n_layers_init', 0) > 0 or o.get('n_layers_out', 1) != 1:
            raise NotImplementedError('encoder != none, n_layers_init > 0 and n_layers_out > 1 are not supported: the '
                                      'encoder branches are dead in the reference; the extra ReLU readout layers '
                                      '(model_attention.py:697-702) are live code there but switched off by the '
                                      'shipped config (n_layers_out=1, config.py) and outside SURVEY 8a')
        params = OrderedDict()
        params['Wemb'] = common.norm_weight(o['n_words'], o['dim_word'])
        self.param_init_fflayer(o, params, 'ff_state', o['ctxg_dim'], o['dim'])
        self.param_init_fflayer(o, params, 'ff_memory', o['ctxg_dim'], o['dim'])
        if o.get('global_proj'):
            self.param_init_fflayer(o, params, 'ff_global', o['ctxg_dim'], o['dim'])
        self.param_init_fflayer(o, params, 'ff_local', o['ctxl_dim'], o['dim'])
        self.param_init_fflayer(o, params, 'ff_motion', o['ctxm_dim'], o['dim'])
        self.param_init_lstm_cond(o, params, o['dim_word'], o['dim'], prefix='decoder')
        self.param_init_fflayer(o, params, 'ff_logit_lstm', o['dim'], o['dim_word'])
        if o['ctx2out']:
            self.param_init_fflayer(o, params, 'ff_logit_ctxglm', o.get('ctxglm_dim', o['dim']), o['dim_word'])
        self.param_init_fflayer(o, params, 'ff_logit', o['dim_word'], o['n_words'])
        return params

    def init_tparams(self, params, force_cpu=False):
        tp = TParams()
        for k, v in params.items():
            tp[k] = Shared(numpy.asarray(v, 'float32'), name=k, owner=tp)
        return tp

    def load_params(self, path, params):
        pp = numpy.load(path)
        for k in params:
            if k not in pp:
                raise Warning('%s is not in the archive' % k)
            params[k] = pp[k]
        return params

    # ---- graph builders ---------------------------------------------------------
    def build_model(self, tparams, options):
        """model_attention.py:583-717.  Returns the reference's 16-tuple; the symbolic
        variables are opaque handles, use ``function(inps, -cost)`` to obtain
        f_log_probs."""
        validate_options(options)
        use_noise = Shared(numpy.float32(0.), name='use_noise')
        model = _Model(tparams, dict(options), use_noise, common.rng_seed)
        trng = model
        hs = [Handle(n, model) for n in BATCH_NAMES]
        alphas = [Handle(n, model) for n in ('alphals', 'alphags', 'alphams', 'alphalts')]
        cost = Handle('cost', model)
        extra = [Handle('probs', model)]
        self._model = model
        return tuple([trng, use_noise] + hs + alphas + [cost, extra])

    def build_sampler(self, tparams, options, use_noise, trng, mode=None):
        """model_attention.py:719-850 -> (f_init, f_next) with numpy in / numpy out."""
        model = trng if isinstance(trng, _Model) else _Model(tparams, dict(options), use_noise, common.rng_seed)
        f32 = torch.float32
        cache = {}
        # the clip's context blocks stay cached in a workspace between f_next calls: a workspace of this
        # sampler alone (no other user of the engine with the same shape may rewrite it)
        tag = ('sampler', id(cache))

        def fingerprint(*arrs):
            fp = []
            for a in arrs:
                a = numpy.asarray(a)
                flat = a.reshape(-1)
                step = max(1, flat.shape[0] // 997)
                fp.append((a.shape, a.__array_interface__['data'][0], float(flat[::step].sum())))
            return tuple(fp)

        def clip_context(ctxg, ctxg_mask, ctxl, ctxm):
            key = fingerprint(ctxg, ctxg_mask, ctxl, ctxm)
            eng = model.engine()
            if cache.get('key') != key or cache.get('version') != model.tparams.version:
                gd = eng.to_device(numpy.asarray(ctxg, 'float32')[None], f32)
                gm = eng.to_device(numpy.asarray(ctxg_mask, 'float32')[None], f32)
                ld = eng.to_device(numpy.asarray(ctxl, 'float32')[None], f32)
                md = eng.to_device(numpy.asarray(ctxm, 'float32')[None], f32)
                cache['max_rows'] = 16
                ws, d = eng.precompute(gd, gm, ld, md, rows=cache['max_rows'], ws_tag=tag)
                cache.update(key=key, version=model.tparams.version, ws=ws, d=d, feats=(gd, gm, ld, md))
            return eng, cache

        def f_init(ctxg, ctxg_mask):
            eng = model.engine()
            gd = eng.to_device(numpy.asarray(ctxg, 'float32')[None], f32)
            gm = eng.to_device(numpy.asarray(ctxg_mask, 'float32')[None], f32)
            h0, c0 = eng.init_state(gd, gm)
            return [numpy.asarray(ctxg, 'float32'), h0.cpu().numpy()[0], c0.cpu().numpy()[0]]

        def f_next(x, ctxg, ctxg_mask, ctxl, ctxl_mask, ctxm, ctxm_mask, h, c):
            eng, cc = clip_context(ctxg, ctxg_mask, ctxl, ctxm)
            k = int(numpy.asarray(x).shape[0])
            if k > cc['max_rows']:
                cc['max_rows'] = k
                cc['ws'], cc['d'] = eng.precompute(*cc['feats'], rows=k, ws_tag=tag)
            xd = eng.to_device(numpy.asarray(x, 'int64'), torch.int64)
            hd = eng.to_device(numpy.asarray(h, 'float32').reshape(k, -1), f32)
            cd = eng.to_device(numpy.asarray(c, 'float32').reshape(k, -1), f32)
            rc = torch.zeros(k, dtype=torch.int32, device=eng.device)
            probs, h2, c2 = eng.step(cc['ws'], cc['d'], xd, hd, cd, row_clip=rc)
            p = probs.cpu().numpy()
            # next_sample (:841) is an MRG multinomial draw that every shipped caller discards
            # (stochastic=False everywhere): the mode is returned unless a caller switches
            # f_next.draw on (gen_sample(stochastic=True) does), then it is a draw from p with numpy's
            # generator (same distribution, not Theano's MRG stream).
            nxt = draw_tokens(p, f_next.rng) if f_next.draw else p.argmax(1)
            return [p, nxt, h2.cpu().numpy(), c2.cpu().numpy()]

        f_next.draw = False
        f_next.rng = numpy.random.RandomState(common.rng_seed)
        return f_init, f_next

    # ---- search (model_attention.py:852-994) ---------------------------------------
    def gen_sample(self, tparams, f_init, f_next, ctxg_0, ctxg_mask, ctxl_0, ctxl_mask, ctxm_0, ctxm_mask,
                   options, trng=None, k=1, maxlen=30, stochastic=False, restrict_voc=False):
        """Beam search (greedy for k=1) for ONE clip; host bookkeeping as in the
        reference: un-normalised cumulative -log p, the k - dead_k best of the flattened
        (live, V) table, hypotheses retire on token 0, survivors appended at the end."""
        if stochastic:
            return self._sample_stochastic(f_init, f_next, ctxg_0, ctxg_mask, ctxl_0, ctxl_mask, ctxm_0, ctxm_mask,
                                           maxlen)
        done, done_scores = [], []
        live = [[]]
        live_scores = numpy.zeros(1, 'float32')
        n_dead = 0
        r = f_init(ctxg_0, ctxg_mask)
        ctxg_0 = r[0]
        state = r[1].reshape(1, -1)
        memory = r[2].reshape(1, -1)
        words = -1 * numpy.ones((1,), 'int64')
        for _ in range(maxlen):
            probs, _, state, memory = f_next(words, ctxg_0, ctxg_mask, ctxl_0, ctxl_mask, ctxm_0, ctxm_mask,
                                             state, memory)
            V = probs.shape[1]
            table = (live_scores[:, None] - numpy.log(probs)).flatten()
            best = table.argsort()[:(k - n_dead)]
            src, wrd = best // V, best % V
            cand = [(live[s] + [int(w)], numpy.float32(table[b]), state[s], memory[s])
                    for s, w, b in zip(src, wrd, best)]
            live, ls, st, me = [], [], [], []
            for hyp, sc, hs, ms in cand:
                if hyp[-1] == 0:
                    done.append(hyp)
                    done_scores.append(sc)
                    n_dead += 1
                else:
                    live.append(hyp)
                    ls.append(sc)
                    st.append(hs)
                    me.append(ms)
            live_scores = numpy.array(ls, 'float32')
            if len(live) < 1 or n_dead >= k:
                break
            words = numpy.array([h[-1] for h in live], 'int64')
            state = numpy.array(st)
            memory = numpy.array(me)
        for hyp, sc in zip(live, live_scores):
            done.append(hyp)
            done_scores.append(sc)
        return done, done_scores, state, memory

    def _sample_stochastic(self, f_init, f_next, ctxg_0, ctxg_mask, ctxl_0, ctxl_mask, ctxm_0, ctxm_mask, maxlen):
        """The stochastic branch of gen_sample (model_attention.py:914-918): follow f_next's own
        next_sample, one hypothesis; `sample` is a flat list of word ids and `sample_score` the SUM OF THE
        PROBABILITIES of the drawn words (as the reference computes it, not a log-likelihood)."""
        sample, score = [], 0.0
        r = f_init(ctxg_0, ctxg_mask)
        ctxg_0, state, memory = r[0], r[1].reshape(1, -1), r[2].reshape(1, -1)
        words = -1 * numpy.ones((1,), 'int64')
        had = getattr(f_next, 'draw', None)
        if had is not None:
            f_next.draw = True
        try:
            for _ in range(maxlen):
                probs, words, state, memory = f_next(words, ctxg_0, ctxg_mask, ctxl_0, ctxl_mask, ctxm_0, ctxm_mask,
                                                     state, memory)
                words = numpy.asarray(words, 'int64')
                sample.append(int(words[0]))
                score += probs[0, words[0]]
                if words[0] == 0:
                    break
        finally:
            if had is not None:
                f_next.draw = had
        return sample, score, state, memory

    # ---- evaluation loop (model_attention.py:996-1032) -------------------------------
    def pred_probs(self, whichset, f_log_probs, verbose=True, prepare_data=None):
        """Mean NLL and perplexity over a split of ``self.engine`` (the data engine).
        ``prepare_data(engine, tags)`` defaults to the engine's own method when it has one (MemoryEngine) and to
        the module function ``data_engine.prepare_data`` otherwise -- the reference's Movie2Caption engine has
        no such method, prepare_data is a module function there (data_engine.py:258)."""
        eng = self.engine
        tags = getattr(eng, whichset)
        iterator = getattr(eng, 'kf_' + whichset)
        prep = prepare_data or getattr(eng, 'prepare_data', None)
        if prep is None:
            from . import data_engine as _de
            prep = _de.prepare_data
        probs, nll, lens = [], [], []
        n_done, n_samples = 0, sum(len(i) for i in iterator)
        for index in iterator:
            batch = prep(eng, [tags[i] for i in index])
            lp = f_log_probs(*batch)
            lens.extend(numpy.asarray(batch[1]).sum(0).tolist())
            nll.extend((-1 * lp).tolist())
            probs.extend(lp.tolist())
            n_done += len(index)
            if verbose:
                sys.stdout.write('\rComputing LL on %d/%d examples' % (n_done, n_samples))
                sys.stdout.flush()
        perp = 2 ** (numpy.sum(nll) / numpy.sum(lens) / numpy.log(2))
        return -1 * numpy.mean(probs), perp

    # ---- training objective (model_attention.py:1129-1147), the first output of f_grad_shared ----
    def train_cost(self, tparams, options, batch, alpha_c=0., decay_c=0.):
        """cost = mean_b(-f_log_probs) + decay_c * sum_params sum(p^2) + alpha_c * sum over the four attentions
        of ((1 - alphas.sum(0))**2).sum(0).mean(), with the current use_noise setting (dropout masks drawn on the
        device when it is 1).  The gradients of this cost (the rest of f_grad_shared) are train.Trainer's."""
        import ctypes as C
        model = getattr(self, '_model', None)
        if model is None or model.tparams is not tparams:
            raise RuntimeError('call build_model(tparams, options) first')
        lp, alphas, _ = model.run_teacher(batch, want_alphas=alpha_c > 0.)
        cost = float((-lp.cpu().numpy().astype('float64')).mean())                       # :1129
        if decay_c > 0.:                                                                  # :1130-1136
            cost += float(decay_c) * sum(float((numpy.asarray(v.get_value(), 'float64') ** 2).sum())
                                         for v in tparams.values())
        if alpha_c > 0.:                                                                  # :1138-1147
            eng = model.engine()
            lib = eng.lib
            scratch = torch.empty((int(lib.stat_clip_scratch_bytes()) + 3) // 4, dtype=torch.float32, device=eng.device)
            out = torch.zeros(4, dtype=torch.float32, device=eng.device)
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for i, a in enumerate(alphas):
                L, rows = a.shape[0], a.shape[1]
                n = int(a.numel() // (L * rows))
                from ._lib import check
                check(lib.stat_alpha_coverage(C.c_void_p(a.data_ptr()), L, rows, n, C.c_void_p(scratch.data_ptr()),
                                              C.c_void_p(out[i:].data_ptr()), stream))
            cost += float(alpha_c) * float(out.cpu().numpy().astype('float64').sum())
        return cost

    # ---- batched extension (not in the reference): beam search for B clips at once ----
    def beam_batch(self, tparams, options, ctxg, mask_ctxg, ctxl, ctxm, k=5, maxlen=30, use_graph=False):
        """gen_sample(k) for every clip of a batch in one device pass.  Host arrays
        in; returns per clip (sample, sample_score) exactly as gen_sample does: lists of word-id
        lists and of cumulative -log p, retired hypotheses first, then the survivors."""
        model = getattr(self, '_model', None)
        if model is None or model.tparams is not tparams:
            model = _Model(tparams, dict(options), Shared(numpy.float32(0.)), common.rng_seed)
            self._model = model
        eng = model.engine()
        f32 = torch.float32
        gd = eng.to_device(ctxg, f32, 'ctxg')
        gm = eng.to_device(mask_ctxg, f32, 'mask_ctxg')
        ld = eng.to_device(ctxl, f32, 'ctxl')
        md = eng.to_device(ctxm, f32, 'ctxm')
        tokens, lengths, scores, count = [t.cpu().numpy() for t in eng.beam_captions(gd, gm, ld, md, k, maxlen, use_graph=use_graph)]
        out = []
        for b in range(tokens.shape[0]):
            n = int(count[b])
            out.append(([tokens[b, j, :lengths[b, j]].tolist() for j in range(n)],
                        [numpy.float32(scores[b, j]) for j in range(n)]))
        return out

    # ---- batched extension (not in the reference): B clips at once -----------------
    def greedy_batch(self, tparams, options, ctxg, mask_ctxg, ctxl, ctxm, maxlen=30, use_graph=True):
        """gen_sample(k=1) for every clip of a batch in one device pass.  Host arrays in,
        (tokens (B,maxlen) int64 with -1 after the eos, lengths, scores) numpy out."""
        model = getattr(self, '_model', None)
        if model is None or model.tparams is not tparams:
            model = _Model(tparams, dict(options), Shared(numpy.float32(0.)), common.rng_seed)
            self._model = model
        eng = model.engine()
        f32 = torch.float32
        gd = eng.to_device(ctxg, f32, 'ctxg')
        gm = eng.to_device(mask_ctxg, f32, 'mask_ctxg')
        ld = eng.to_device(ctxl, f32, 'ctxl')
        md = eng.to_device(ctxm, f32, 'ctxm')
        tokens, lengths, scores = eng.greedy_captions(gd, gm, ld, md, maxlen, use_graph=use_graph)
        return tokens.cpu().numpy(), lengths.cpu().numpy(), scores.cpu().numpy()
