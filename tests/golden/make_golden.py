"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/model_attention.py via oracle/ref_exec.py; see that file for
how).  Run in the authoring container only:

    python tests/golden/make_golden.py

Each fixture stores the options, the parameters, the prepare_data-shaped
inputs and the outputs of the reference callables f_log_probs / extra /
f_init / f_next / gen_sample.  tests/test_oracle_golden.py replays them
against oracle/stat_oracle.py (no reference needed at test time).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_exec, stat_oracle as so                      # noqa: E402
from video_description_with_spatial_temporal_attention_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_case(name, opt_kw, B, T, R, L, params_kind, seed, ks=(1, 3, 5), maxlen=8,
             eos_bias=0.0, nonneg=False):
    o = so.default_options(**opt_kw)
    if params_kind == 'init':
        params = None
    else:
        params = so.trained_like_params(o, seed=seed)
        params['ff_logit_b'] = params['ff_logit_b'].copy()
        params['ff_logit_b'][0] += eos_bias
    rm = ref_exec.RefModel(o, params=params)
    params = rm.params
    batch = synthetic.make_batch(o, B=B, T=T, R=R, L=L, seed=seed, zero_tail=True,
                                 nonneg=nonneg)
    x, m, ctxg, mg, ctxl, ml, ctxm, mm = batch
    d = {'options': np.array(json.dumps(o))}
    for k, v in params.items():
        d['param:' + k] = np.asarray(v)
    d['param_order'] = np.array(json.dumps(list(params.keys())))
    for n, a in zip(('x', 'mask', 'ctxg', 'ctxg_mask', 'ctxl', 'ctxl_mask', 'ctxm',
                     'ctxm_mask'), batch):
        d['in:' + n] = a
    d['out:f_log_probs'] = np.asarray(rm.f_log_probs(*batch))
    probs, als, ags, ams, alts = rm.f_extra(*batch)
    d['out:probs'], d['out:alphals'], d['out:alphags'] = probs, als, ags
    d['out:alphams'], d['out:alphalts'] = ams, alts
    # sampler: f_init and two chained f_next calls on clip 0
    r = rm.f_init(ctxg[0], mg[0])
    d['out:f_init_h0'], d['out:f_init_c0'] = r[1], r[2]
    nw = -np.ones((1,), 'int64')
    r1 = rm.f_next(nw, ctxg[0], mg[0], ctxl[0], ml[0], ctxm[0], mm[0],
                   r[1][None], r[2][None])
    d['out:f_next0_probs'], d['out:f_next0_h'], d['out:f_next0_c'] = r1[0], r1[2], r1[3]
    nw2 = np.array([int(r1[0][0].argmax()), 3], 'int64')
    h2 = np.concatenate([r1[2], r1[2] * 0.5]); c2 = np.concatenate([r1[3], r1[3] - 0.1])
    r2 = rm.f_next(nw2, ctxg[0], mg[0], ctxl[0], ml[0], ctxm[0], mm[0], h2, c2)
    d['in:f_next1_x'], d['in:f_next1_h'], d['in:f_next1_c'] = nw2, h2, c2
    d['out:f_next1_probs'], d['out:f_next1_h'], d['out:f_next1_c'] = r2[0], r2[2], r2[3]
    # beam search per clip
    info = {}
    for k in ks:
        for b in range(B):
            s, sc, _, _ = rm.gen_sample(ctxg[b], mg[b], ctxl[b], ml[b], ctxm[b], mm[b],
                                        k, maxlen)
            d['out:gen_k%d_b%d_scores' % (k, b)] = np.asarray(sc, 'float32')
            flat = -np.ones((len(s), maxlen), 'int64')
            for i, hyp in enumerate(s):
                flat[i, :len(hyp)] = hyp
            d['out:gen_k%d_b%d_tokens' % (k, b)] = flat
            info[(k, b)] = [len(h) for h in s]
    d['gen_ks'] = np.asarray(ks)
    d['gen_maxlen'] = np.asarray(maxlen)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
    print(name, 'f_log_probs', d['out:f_log_probs'], 'hyp lengths', info)


if __name__ == '__main__':
    assert ref_exec.mounted(), 'needs /root/reference'
    only = sys.argv[1:]            # optional: fixture names to (re)generate
    _run = run_case

    def run_case(name, *a, **kw):  # noqa: F811
        if not only or name in only:
            _run(name, *a, **kw)
    tiny = dict(dim_word=8, dim=8, ctxg_dim=8, ctxl_dim=12, ctxm_dim=10, n_words=11)
    run_case('ref_tiny_trained', tiny, B=3, T=4, R=2, L=5, params_kind='trained',
             seed=11, eos_bias=1.5)
    run_case('ref_tiny_init', tiny, B=3, T=4, R=2, L=5, params_kind='init', seed=12)
    run_case('ref_tiny_noflags', dict(tiny, selector=False, prev2out=False, ctx2out=False),
             B=2, T=3, R=2, L=4, params_kind='trained', seed=13, eos_bias=1.0)
    mid = dict(dim_word=24, dim=32, ctxg_dim=32, ctxl_dim=40, ctxm_dim=36, n_words=50)
    run_case('ref_mid_trained', mid, B=4, T=6, R=3, L=7, params_kind='trained',
             seed=14, eos_bias=2.5, maxlen=10, nonneg=True)
    # decision D1 (ctxg_dim != dim): the reference's own commented-out ff_global lines
    # (model_attention.py:553-554, 661-662, 780-781), un-commented in memory by ref_exec
    run_case('ref_tiny_globalproj', dict(tiny, ctxg_dim=20, global_proj=True), B=3, T=4, R=2, L=5,
             params_kind='trained', seed=21, eos_bias=1.5)
    run_case('ref_tiny_globalproj_init', dict(tiny, ctxg_dim=20, global_proj=True), B=2, T=4, R=2, L=5,
             params_kind='init', seed=22)
    run_case('ref_mid_globalproj', dict(mid, ctxg_dim=48, global_proj=True), B=4, T=6, R=3, L=7,
             params_kind='trained', seed=23, eos_bias=0.8, maxlen=10, nonneg=True)
