"""CPU oracle for the spatial-temporal-attention caption decoder.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the CPU reference being timed.

What it is: a plain-numpy restatement of the reference's Theano graph for the
hot path, written from the equations, with every
intermediate exposed so a kernel bug can be localised.  ``dtype`` selects fp32
(the reference's floatX) or fp64 (truth, to bound the fp32 oracle's own error).

Pinning status: the reference holds no golden vectors for this path and Theano
is not installable here, so the reference itself cannot be imported.  The
oracle is instead pinned by *executing the reference's own source text*
(``/root/reference/model_attention.py``, py2->py3 token fixes applied in memory,
on a small eager stand-in for the handful of Theano ops it uses) and comparing
outputs: see ``oracle/ref_exec.py`` and ``tests/golden/``.  Without those
fixtures this oracle would be "parity unpinned".

Reference line citations are ``model_attention.py:<line>`` unless prefixed.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

# ----------------------------------------------------------------------------
# options
# ----------------------------------------------------------------------------

def default_options(**kw):
    """Option keys consumed by the hot path (config.py:17-49, train() kwargs
    model_attention.py:1034-1078).  ``global_proj`` is decision D1: it enables
    the layer the reference left commented out (:553-554, :661-662, :780-781)
    so that ctxg_dim may differ from dim."""
    o = dict(dim_word=512, dim=512, ctxg_dim=512, ctxl_dim=4096, ctxm_dim=4096,
             ctxglm_dim=512, n_words=12594, selector=True, prev2out=True,
             ctx2out=True, use_dropout=True, n_layers_out=1, n_layers_init=0,
             encoder='none', global_proj=False)
    o.update(kw)
    if not o['global_proj'] and o['ctxg_dim'] != o['dim']:
        raise ValueError('reference graph requires ctxg_dim == dim '
                         '(SURVEY F3); set global_proj=True for ctxg_dim != dim')
    o['ctxglm_dim'] = o['dim']
    return o


# ----------------------------------------------------------------------------
# a1: parameters  (model_attention.py:518-581, :180-282, :80-87; common.py:110-134)
# ----------------------------------------------------------------------------

class _Init:
    """The reference draws every tensor from one module-global
    RandomState(1234) (common.py:16-25) in creation order."""

    def __init__(self, seed=1234):
        self.rng = np.random.RandomState(seed)

    def ortho(self, n):                                   # common.py:110-122
        u, _, _ = np.linalg.svd(self.rng.randn(n, n))
        return u.astype('float32')

    def norm(self, nin, nout=None, scale=0.01, ortho=True):  # common.py:124-134
        if nout is None:
            nout = nin
        if nout == nin and ortho:
            return self.ortho(nin)
        return (scale * self.rng.randn(nin, nout)).astype('float32')


def init_params(options, seed=1234):
    o = options
    g = _Init(seed)
    H, E = o['dim'], o['dim_word']
    p = OrderedDict()

    def ff(prefix, nin, nout):                            # :80-87
        p[prefix + '_W'] = g.norm(nin, nout, scale=0.01)
        p[prefix + '_b'] = np.zeros((nout,), 'float32')

    p['Wemb'] = g.norm(o['n_words'], E)                   # :522
    ff('ff_state', o['ctxg_dim'], H)                      # :549-550
    ff('ff_memory', o['ctxg_dim'], H)                     # :551-552
    if o.get('global_proj'):
        ff('ff_global', o['ctxg_dim'], H)                 # :553-554 (D1)
    ff('ff_local', o['ctxl_dim'], H)                      # :556-557
    ff('ff_motion', o['ctxm_dim'], H)                     # :558-559
    d = 'decoder_'                                        # :561-563, :180-282
    p[d + 'W'] = np.concatenate([g.norm(E, H) for _ in range(4)], axis=1)
    p[d + 'U'] = np.concatenate([g.ortho(H) for _ in range(4)], axis=1)
    p[d + 'b'] = np.zeros((4 * H,), 'float32')
    p[d + 'Wc'] = g.norm(H, 4 * H)
    p[d + 'Wcg_att'] = g.norm(H, ortho=False)
    p[d + 'Wcm_att'] = g.norm(H, ortho=False)
    p[d + 'Wclt_att'] = g.norm(H, ortho=False)
    p[d + 'Wdg_att'] = g.norm(H, H)
    p[d + 'Wdm_att'] = g.norm(H, H)
    p[d + 'Wdlt_att'] = g.norm(H, H)
    p[d + 'bg_att'] = np.zeros((H,), 'float32')
    p[d + 'bm_att'] = np.zeros((H,), 'float32')
    p[d + 'blt_att'] = np.zeros((H,), 'float32')
    p[d + 'Wcl_att'] = g.norm(H, ortho=False)
    p[d + 'Wdl_att'] = g.norm(H, H)
    p[d + 'bl_att'] = np.zeros((H,), 'float32')
    for nm in ('g', 'm', 'lt', 'l'):                      # :255-274
        p[d + 'U%s_att' % nm] = g.norm(H, 1)
        p[d + 'c%s_att' % nm] = np.zeros((1,), 'float32')
    if o['selector']:                                     # :276-281
        p[d + 'W_sel'] = g.norm(H, 1)
        p[d + 'b_sel'] = np.float32(0.)
    ff('ff_logit_lstm', H, E)                             # :566-568
    if o['ctx2out']:
        ff('ff_logit_ctxglm', o['ctxglm_dim'], E)         # :569-572
    if o['n_layers_out'] > 1:
        for l in range(1, o['n_layers_out']):
            ff('ff_logit_h%d' % l, E, E)                  # :573-577
    ff('ff_logit', E, o['n_words'])                       # :578-580
    return p


def trained_like_params(options, seed=7):
    """Same key set / shapes as init_params but with magnitudes resembling a
    trained model, so the softmaxes are far from uniform and parity is not
    trivially easy.  Not from the reference: test input only."""
    p = init_params(options, seed=1234)
    r = np.random.RandomState(seed)
    H = options['dim']
    out = OrderedDict()
    for k, v in p.items():
        v = np.asarray(v, 'float32')
        if v.ndim == 2 and v.shape[1] == 1:               # score vectors (H,1)
            out[k] = (r.randn(*v.shape) * (2.0 / np.sqrt(H))).astype('float32')
        elif v.ndim == 2:
            nin = v.shape[0]
            gain = 1.0
            if k == 'Wemb':
                out[k] = (r.randn(*v.shape) * 0.5).astype('float32')
                continue
            if k == 'ff_logit_W':
                gain = 3.0
            if k.startswith('ff_local') or k.startswith('ff_motion') or \
               k.startswith('ff_global') or k.startswith('ff_state') or \
               k.startswith('ff_memory'):
                gain = 2.0                                 # inputs have std 0.5
            out[k] = (r.randn(*v.shape) * (gain / np.sqrt(nin))).astype('float32')
        elif v.ndim == 1:
            out[k] = (r.randn(*v.shape) * 0.1).astype('float32')
        else:                                              # b_sel scalar
            out[k] = np.float32(r.randn() * 0.1)
    return out


# ----------------------------------------------------------------------------
# elementwise helpers (Theano semantics: max-subtracted softmax)
# ----------------------------------------------------------------------------

def _softmax(a, axis=-1):
    m = a.max(axis=axis, keepdims=True)
    e = np.exp(a - m)
    return e / e.sum(axis=axis, keepdims=True)


def _sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


def _cast(params, dtype):
    return {k: np.asarray(v, dtype) for k, v in params.items()}


# ----------------------------------------------------------------------------
# a2 + a3: prologue  (build_model :609-667, lstm_cond_layer :322-335)
# ----------------------------------------------------------------------------

def init_state(P, o, ctxg, mask_ctxg):
    """P2, P3: mean over *all* frames divided by the count of non-zero frames
    (:618, :649), then ff_state / ff_memory with tanh (:657-660)."""
    counts = mask_ctxg.sum(-1)[:, None]
    gbar = ctxg.sum(1) / counts
    h0 = np.tanh(gbar @ P['ff_state_W'] + P['ff_state_b'])
    c0 = np.tanh(gbar @ P['ff_memory_W'] + P['ff_memory_b'])
    return gbar, h0, c0


def project_contexts(P, o, ctxg, ctxl, ctxm):
    """P4, P5.  ctxg (B,T,Dg) ctxl (B,T,R,Dr) ctxm (B,T,Dm)."""
    if o.get('global_proj'):
        G = np.tanh(ctxg @ P['ff_global_W'] + P['ff_global_b'])   # D1
    else:
        G = ctxg                                                   # :646, :671
    Lc = np.tanh(ctxl @ P['ff_local_W'] + P['ff_local_b'])         # :664-665
    M = np.tanh(ctxm @ P['ff_motion_W'] + P['ff_motion_b'])        # :666-667
    pG = G @ P['decoder_Wcg_att'] + P['decoder_bg_att']            # :322
    pL = Lc @ P['decoder_Wcl_att'] + P['decoder_bl_att']           # :324
    pM = M @ P['decoder_Wcm_att'] + P['decoder_bm_att']            # :326
    return dict(G=G, Lc=Lc, M=M, pG=pG, pL=pL, pM=pM)


# ----------------------------------------------------------------------------
# a4: one step of the recurrent cell (_step :366-459)
# ----------------------------------------------------------------------------

def cell_step(P, o, blk, m, x, h_, c_, dp):
    """m (B,) mask, x (B,4H) = emb.W + b, h_/c_ (B,H), dp (B,3H) dropout
    factors (0.5 at eval, :469-477).  blk from project_contexts; leading dim of
    the blocks may be 1 (sampler: one clip shared by k hypotheses, :786-788)."""
    H = h_.shape[1]
    d = 'decoder_'
    # S1-S3 spatial attention over the R regions of each frame (:371-383)
    sl = h_ @ P[d + 'Wdl_att']
    aL = np.tanh(blk['pL'] + sl[:, None, None, :]) @ P[d + 'Ul_att'] + P[d + 'cl_att']
    aL = aL[..., 0]                                            # (B,T,R)
    alphaL = _softmax(aL, axis=-1)
    cL = (blk['Lc'] * alphaL[..., None]).sum(2)                # (B,T,H)
    # S4 temporal attention on global features (:389-399)
    aG = (np.tanh(blk['pG'] + (h_ @ P[d + 'Wdg_att'])[:, None, :]) @ P[d + 'Ug_att']
          + P[d + 'cg_att'])[..., 0]
    alphaG = _softmax(aG, axis=-1)
    cG = (blk['G'] * alphaG[..., None]).sum(1)
    # S5 temporal attention on motion features (:402-412)
    aM = (np.tanh(blk['pM'] + (h_ @ P[d + 'Wdm_att'])[:, None, :]) @ P[d + 'Um_att']
          + P[d + 'cm_att'])[..., 0]
    alphaM = _softmax(aM, axis=-1)
    cM = (blk['M'] * alphaM[..., None]).sum(1)
    # S6-S7 temporal attention on the spatially attended local context (:415-426)
    pLT = cL @ P[d + 'Wclt_att'] + P[d + 'blt_att'] + (h_ @ P[d + 'Wdlt_att'])[:, None, :]
    aLT = (np.tanh(pLT) @ P[d + 'Ult_att'] + P[d + 'clt_att'])[..., 0]
    alphaLT = _softmax(aLT, axis=-1)
    cLT = (cL * alphaLT[..., None]).sum(1)
    # S8-S9 fusion by sum, scalar selector gate (:430-435)
    ctx = cG + cM + cLT
    beta = None
    if o['selector']:
        beta = _sigmoid(h_ @ P[d + 'W_sel'] + P[d + 'b_sel'])[:, 0]
        ctx = beta[:, None] * ctx
    # S10-S13 LSTM gates; dropout factor multiplies i,f,o *pre*-activations
    # (:437-457); g has none.
    pre = h_ @ P[d + 'U'] + x + ctx @ P[d + 'Wc']
    i = _sigmoid(pre[:, 0:H] * dp[:, 0:H])
    f = _sigmoid(pre[:, H:2 * H] * dp[:, H:2 * H])
    og = _sigmoid(pre[:, 2 * H:3 * H] * dp[:, 2 * H:3 * H])
    g = np.tanh(pre[:, 3 * H:4 * H])
    c = f * c_ + i * g
    c = m[:, None] * c + (1. - m)[:, None] * c_
    h = og * np.tanh(c)
    h = m[:, None] * h + (1. - m)[:, None] * h_
    return dict(h=h, c=c, alphaL=alphaL, alphaG=alphaG, alphaM=alphaM,
                alphaLT=alphaLT, cL=cL, cG=cG, cM=cM, cLT=cLT, ctx=ctx,
                beta=beta, pre=pre, aL=aL, aG=aG, aM=aM, aLT=aLT)


# ----------------------------------------------------------------------------
# a6: readout (:684-709 / :817-840)
# ----------------------------------------------------------------------------

def readout_logits(P, o, h, emb, ctx, dp_h=0.5, dp_z=0.5):
    z = (h * dp_h) @ P['ff_logit_lstm_W'] + P['ff_logit_lstm_b']
    if o['prev2out']:
        z = z + emb
    if o['ctx2out']:
        z = z + ctx @ P['ff_logit_ctxglm_W'] + P['ff_logit_ctxglm_b']
    z = np.tanh(z) * dp_z
    for l in range(1, o['n_layers_out']):
        z = np.maximum(0., z @ P['ff_logit_h%d_W' % l] + P['ff_logit_h%d_b' % l]) * dp_z
    return z @ P['ff_logit_W'] + P['ff_logit_b']


# ----------------------------------------------------------------------------
# f_log_probs: teacher-forced forward (build_model :583-717, :1126)
# ----------------------------------------------------------------------------

def forward_teacher(params, options, x, mask, ctxg, mask_ctxg, ctxl, mask_ctxl,
                    ctxm, mask_ctxm, dtype=np.float32, return_all=False,
                    dp_gates=None, dp_h=None, dp_z=None):
    """Returns f_log_probs (B,) = sum_t mask*log(p[x]+1e-8), use_noise=0 unless
    explicit dropout factors are given (D2).  mask_ctxl / mask_ctxm are
    accepted and ignored exactly like the reference."""
    o = options
    P = _cast(params, dtype)
    x = np.asarray(x)
    mask = np.asarray(mask, dtype)
    ctxg, ctxl, ctxm = (np.asarray(a, dtype) for a in (ctxg, ctxl, ctxm))
    mask_ctxg = np.asarray(mask_ctxg, dtype)
    L, B = x.shape
    H = o['dim']
    emb = P['Wemb'][x.reshape(-1)].reshape(L, B, -1)              # :613
    emb = np.concatenate([np.zeros_like(emb[:1]), emb[:-1]], 0)  # :615-617
    gbar, h, c = init_state(P, o, ctxg, mask_ctxg)
    blk = project_contexts(P, o, ctxg, ctxl, ctxm)
    X = emb @ P['decoder_W'] + P['decoder_b']                     # :334-335
    half = np.asarray(0.5, dtype)
    steps = []
    logp = np.zeros((B,), dtype)
    for t in range(L):
        dp = dp_gates[t] if dp_gates is not None else np.full((B, 3 * H), half, dtype)
        s = cell_step(P, o, blk, mask[t], X[t], h, c, dp)
        h, c = s['h'], s['c']
        logits = readout_logits(P, o, h, emb[t], s['ctx'],
                                dp_h[t] if dp_h is not None else half,
                                dp_z[t] if dp_z is not None else half)
        p = _softmax(logits, axis=-1)                             # :708-709
        tok = p[np.arange(B), x[t]]
        logp = logp + mask[t] * np.log(tok + np.asarray(1e-8, dtype))   # :712-715
        if return_all:
            s['logits'] = logits
            s['probs'] = p
            steps.append(s)
    if return_all:
        return logp, dict(h0=init_state(P, o, ctxg, mask_ctxg)[1],
                          c0=init_state(P, o, ctxg, mask_ctxg)[2],
                          gbar=gbar, blk=blk, X=X, emb=emb, steps=steps)
    return logp


# ----------------------------------------------------------------------------
# a7, a8: sampler functions (build_sampler :719-850)
# ----------------------------------------------------------------------------

def make_sampler(params, options, dtype=np.float32, hoist=False):
    """Returns (f_init, f_next) with the reference's positional signatures.

    faithful (hoist=False): every f_next call recomputes tanh(ctxl.W_local),
    tanh(ctxm.W_motion) and the three projected blocks from the raw features,
    as the compiled Theano function does (:782-788, :806-814).
    hoist=True caches them per clip (same numbers, used for fast goldens)."""
    o = options
    P = _cast(params, dtype)
    H = o['dim']
    cache = {}

    def f_init(ctxg, ctxg_mask):                                  # :791-795
        ctxg = np.asarray(ctxg, dtype)
        ctxg_mask = np.asarray(ctxg_mask, dtype)
        _, h0, c0 = init_state(P, o, ctxg[None], ctxg_mask[None])
        return [ctxg, h0[0], c0[0]]

    def f_next(x, ctxg, ctxg_mask, ctxl, ctxl_mask, ctxm, ctxm_mask, h, c):
        x = np.asarray(x)
        h = np.asarray(h, dtype)
        c = np.asarray(c, dtype)
        key = (id(ctxg), id(ctxl), id(ctxm))
        if hoist and key in cache:
            blk = cache[key]
        else:
            blk = project_contexts(P, o, np.asarray(ctxg, dtype)[None],
                                   np.asarray(ctxl, dtype)[None],
                                   np.asarray(ctxm, dtype)[None])
            if hoist:
                cache.clear()
                cache[key] = blk
        k = x.shape[0]
        emb = np.where((x < 0)[:, None], np.zeros((1, P['Wemb'].shape[1]), dtype),
                       P['Wemb'][np.maximum(x, 0)])                # :803-804
        X = emb @ P['decoder_W'] + P['decoder_b']
        ones = np.ones((k,), dtype)            # mask is identity in one-step mode
        dp = np.full((k, 3 * H), 0.5, dtype)
        s = cell_step(P, o, blk, ones, X, h, c, dp)
        logits = readout_logits(P, o, s['h'], emb, s['ctx'],
                                np.asarray(0.5, dtype), np.asarray(0.5, dtype))
        probs = _softmax(logits, axis=-1)                          # :840
        # next_sample (:841) is an MRG multinomial draw that every caller
        # discards (stochastic=False everywhere); return argmax.
        return [probs, probs.argmax(1), s['h'], s['c']]

    return f_init, f_next


# ----------------------------------------------------------------------------
# a9: gen_sample  (:852-994)  -- beam search / greedy for one clip
# ----------------------------------------------------------------------------

def gen_sample(f_init, f_next, ctxg, ctxg_mask, ctxl, ctxl_mask, ctxm, ctxm_mask,
               k=1, maxlen=30):
    """Deterministic branch only (stochastic=False is the only one any caller
    uses).  Semantics kept: un-normalised cumulative -log p scores (:921),
    candidates = the k-dead_k smallest of the flattened (live,V) score table
    (:923), a hypothesis retires when it emits token 0 (:958-962), loop stops
    when no live hypothesis remains or dead>=k (:974-977), survivors are
    appended after the loop (:987-992)."""
    sample, sample_score = [], []
    live_k, dead_k = 1, 0
    hyp_samples = [[]]
    hyp_scores = np.zeros(1, 'float32')
    r = f_init(ctxg, ctxg_mask)
    ctxg = r[0]
    next_state = r[1].reshape(1, -1)
    next_memory = r[2].reshape(1, -1)
    next_w = -1 * np.ones((1,), 'int64')                          # :893
    for _ in range(maxlen):
        next_p, _, next_state, next_memory = f_next(
            next_w, ctxg, ctxg_mask, ctxl, ctxl_mask, ctxm, ctxm_mask,
            next_state, next_memory)
        cand = hyp_scores[:, None] - np.log(next_p)               # :921
        flat = cand.flatten()
        ranks = flat.argsort()[:(k - dead_k)]                     # :923
        V = next_p.shape[1]
        ti = ranks // V                                           # :926 (py2 int /)
        wi = ranks % V
        costs = flat[ranks]
        new_samples = [hyp_samples[t] + [int(w)] for t, w in zip(ti, wi)]
        new_scores = np.asarray(costs, 'float32')                 # :931, :941
        new_states = [next_state[t].copy() for t in ti]
        new_memories = [next_memory[t].copy() for t in ti]
        hyp_samples, hs, hst, hm = [], [], [], []
        for idx in range(len(new_samples)):
            if new_samples[idx][-1] == 0:
                sample.append(new_samples[idx])
                sample_score.append(new_scores[idx])
                dead_k += 1
            else:
                hyp_samples.append(new_samples[idx])
                hs.append(new_scores[idx])
                hst.append(new_states[idx])
                hm.append(new_memories[idx])
        hyp_scores = np.array(hs)
        live_k = len(hyp_samples)
        if live_k < 1 or dead_k >= k:
            break
        next_w = np.array([w[-1] for w in hyp_samples])
        next_state = np.array(hst)
        next_memory = np.array(hm)
    if live_k > 0:
        for idx in range(live_k):
            sample.append(hyp_samples[idx])
            sample_score.append(hyp_scores[idx])
    return sample, sample_score, next_state, next_memory


# ----------------------------------------------------------------------------
# batched greedy decode: what gen_sample(k=1) computes, for B clips at once.
# Used as the oracle for the device greedy loop (config 2).
# ----------------------------------------------------------------------------

def greedy_decode_batch(params, options, ctxg, mask_ctxg, ctxl, ctxm, maxlen,
                        dtype=np.float32, return_all=False):
    """Per clip identical to gen_sample(k=1): token_t = argmax p_t, score +=
    -log p_t[token_t], stop after emitting 0.  Returns tokens (B,maxlen) int64
    padded with -1 after the eos, lengths (B,), scores (B,) and, optionally,
    the top-2 logit margin of every live step (for tie diagnostics)."""
    o = options
    P = _cast(params, dtype)
    ctxg, ctxl, ctxm = (np.asarray(a, dtype) for a in (ctxg, ctxl, ctxm))
    B = ctxg.shape[0]
    H = o['dim']
    _, h, c = init_state(P, o, ctxg, np.asarray(mask_ctxg, dtype))
    blk = project_contexts(P, o, ctxg, ctxl, ctxm)
    tokens = -np.ones((B, maxlen), 'int64')
    scores = np.zeros((B,), dtype)
    lengths = np.zeros((B,), 'int64')
    alive = np.ones((B,), bool)
    prev = -np.ones((B,), 'int64')
    margins = np.full((B, maxlen), np.inf)
    half = np.asarray(0.5, dtype)
    for t in range(maxlen):
        emb = np.where((prev < 0)[:, None], np.zeros((1, P['Wemb'].shape[1]), dtype),
                       P['Wemb'][np.maximum(prev, 0)])
        X = emb @ P['decoder_W'] + P['decoder_b']
        s = cell_step(P, o, blk, np.ones((B,), dtype), X, h, c,
                      np.full((B, 3 * H), half, dtype))
        logits = readout_logits(P, o, s['h'], emb, s['ctx'], half, half)
        p = _softmax(logits, axis=-1)
        w = (-np.log(p)).argmin(1)
        srt = np.sort(logits, axis=1)
        margins[alive, t] = (srt[:, -1] - srt[:, -2])[alive]
        scores = np.where(alive, scores - np.log(p[np.arange(B), w]), scores)
        tokens[alive, t] = w[alive]
        lengths = np.where(alive, t + 1, lengths)
        # dead clips keep stepping with frozen state; they no longer matter
        h = np.where(alive[:, None], s['h'], h)
        c = np.where(alive[:, None], s['c'], c)
        prev = np.where(alive, w, prev)
        alive = alive & (w != 0)
        if not alive.any():
            break
    if return_all:
        return tokens, lengths, scores, margins
    return tokens, lengths, scores
