"""Gradient oracle for the training step (TEST INFRASTRUCTURE ONLY, same import rules as
stat_oracle.py): the teacher-forced forward of ``stat_oracle.forward_teacher`` restated op for op in
torch (CPU, float64 by default) so that autograd yields what ``tensor.grad(cost, wrt=itemlist(tparams))``
yields in the reference (model_attention.py:1193), for the training cost of model_attention.py:1129-1147:

    cost = mean_b( -sum_t mask * log(p[x] + 1e-8) )
         + decay_c * sum_params sum(p^2)                                            (:1130-1136)
         + alpha_c * sum over the four attentions of ((1 - alphas.sum(0))**2).sum(0).mean()   (:1138-1147)

followed by the global-norm clipping of :1194-1203.  The forward is checked against the numpy oracle
(itself pinned to the reference's source) in tests/test_grad_oracle.py; the gradients against central
differences.  The CUDA backward pass (csrc/backward.cu, stat_grad_shared) is compared with this in
tests/test_gpu_train.py (GPU) and tests/test_backward_emu.py (the same code under the CPU launch emulation).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def forward(P, o, x, mask, ctxg, mask_ctxg, ctxl, ctxm, dp_gates=None, dp_h=None, dp_z=None):
    """torch twin of stat_oracle.forward_teacher (same equations, same order).  P: dict of tensors.
    Returns (logp (B,), alphas dict of stacked (L,...) tensors)."""
    L, B = x.shape
    H = o['dim']
    d = 'decoder_'
    dt = P['Wemb'].dtype
    emb = P['Wemb'][x.reshape(-1)].reshape(L, B, -1)                               # :613
    emb = torch.cat([torch.zeros_like(emb[:1]), emb[:-1]], 0)                       # :615-617
    counts = mask_ctxg.sum(-1)[:, None]
    gbar = ctxg.sum(1) / counts                                                     # :618, :649
    h = torch.tanh(gbar @ P['ff_state_W'] + P['ff_state_b'])                        # :657-660
    c = torch.tanh(gbar @ P['ff_memory_W'] + P['ff_memory_b'])
    if o.get('global_proj'):
        G = torch.tanh(ctxg @ P['ff_global_W'] + P['ff_global_b'])                  # D1
    else:
        G = ctxg
    Lc = torch.tanh(ctxl @ P['ff_local_W'] + P['ff_local_b'])                       # :664-665
    M = torch.tanh(ctxm @ P['ff_motion_W'] + P['ff_motion_b'])                      # :666-667
    pG = G @ P[d + 'Wcg_att'] + P[d + 'bg_att']                                     # :322
    pL = Lc @ P[d + 'Wcl_att'] + P[d + 'bl_att']                                    # :324
    pM = M @ P[d + 'Wcm_att'] + P[d + 'bm_att']                                     # :326
    X = emb @ P[d + 'W'] + P[d + 'b']                                               # :334-335
    half = torch.tensor(0.5, dtype=dt)
    logp = torch.zeros(B, dtype=dt)
    al, ag, am, alt = [], [], [], []
    for t in range(L):
        h_, c_, m = h, c, mask[t]
        dp = dp_gates[t] if dp_gates is not None else torch.full((B, 3 * H), 0.5, dtype=dt)
        # S1-S3 (:371-383)
        sl = h_ @ P[d + 'Wdl_att']
        aL = (torch.tanh(pL + sl[:, None, None, :]) @ P[d + 'Ul_att'] + P[d + 'cl_att'])[..., 0]
        alphaL = torch.softmax(aL, dim=-1)
        cL = (Lc * alphaL[..., None]).sum(2)
        # S4 (:389-399)
        aG = (torch.tanh(pG + (h_ @ P[d + 'Wdg_att'])[:, None, :]) @ P[d + 'Ug_att'] + P[d + 'cg_att'])[..., 0]
        alphaG = torch.softmax(aG, dim=-1)
        cG = (G * alphaG[..., None]).sum(1)
        # S5 (:402-412)
        aM = (torch.tanh(pM + (h_ @ P[d + 'Wdm_att'])[:, None, :]) @ P[d + 'Um_att'] + P[d + 'cm_att'])[..., 0]
        alphaM = torch.softmax(aM, dim=-1)
        cM = (M * alphaM[..., None]).sum(1)
        # S6-S7 (:415-426)
        pLT = cL @ P[d + 'Wclt_att'] + P[d + 'blt_att'] + (h_ @ P[d + 'Wdlt_att'])[:, None, :]
        aLT = (torch.tanh(pLT) @ P[d + 'Ult_att'] + P[d + 'clt_att'])[..., 0]
        alphaLT = torch.softmax(aLT, dim=-1)
        cLT = (cL * alphaLT[..., None]).sum(1)
        # S8-S9 (:430-435)
        ctx = cG + cM + cLT
        if o['selector']:
            beta = torch.sigmoid(h_ @ P[d + 'W_sel'] + P[d + 'b_sel'])[:, 0]
            ctx = beta[:, None] * ctx
        # S10-S13 (:437-457)
        pre = h_ @ P[d + 'U'] + X[t] + ctx @ P[d + 'Wc']
        i = torch.sigmoid(pre[:, 0:H] * dp[:, 0:H])
        f = torch.sigmoid(pre[:, H:2 * H] * dp[:, H:2 * H])
        og = torch.sigmoid(pre[:, 2 * H:3 * H] * dp[:, 2 * H:3 * H])
        g = torch.tanh(pre[:, 3 * H:4 * H])
        c = f * c_ + i * g
        c = m[:, None] * c + (1. - m)[:, None] * c_
        h = og * torch.tanh(c)
        h = m[:, None] * h + (1. - m)[:, None] * h_
        # readout (:684-709)
        z = (h * (dp_h[t] if dp_h is not None else half)) @ P['ff_logit_lstm_W'] + P['ff_logit_lstm_b']
        if o['prev2out']:
            z = z + emb[t]
        if o['ctx2out']:
            z = z + ctx @ P['ff_logit_ctxglm_W'] + P['ff_logit_ctxglm_b']
        z = torch.tanh(z) * (dp_z[t] if dp_z is not None else half)
        logits = z @ P['ff_logit_W'] + P['ff_logit_b']
        p = torch.softmax(logits, dim=-1)
        tok = p[torch.arange(B), x[t]]
        logp = logp + m * torch.log(tok + 1e-8)                                     # :712-715
        al.append(alphaL); ag.append(alphaG); am.append(alphaM); alt.append(alphaLT)
    return logp, dict(alphals=torch.stack(al), alphags=torch.stack(ag), alphams=torch.stack(am),
                      alphalts=torch.stack(alt))


def training_cost(P, o, x, mask, ctxg, mask_ctxg, ctxl, ctxm, alpha_c=0., decay_c=0., **dp):
    """model_attention.py:1129-1147 (alpha_entropy_r is 0 in every shipped config and its branch
    references undefined names, SURVEY App. C: not restated)."""
    logp, al = forward(P, o, x, mask, ctxg, mask_ctxg, ctxl, ctxm, **dp)
    cost = (-logp).mean()                                                           # :1129
    if decay_c > 0.:
        cost = cost + decay_c * sum((v ** 2).sum() for v in P.values())             # :1130-1136
    if alpha_c > 0.:
        for k in ('alphags', 'alphals', 'alphams', 'alphalts'):                     # :1138-1147
            cost = cost + alpha_c * ((1. - al[k].sum(0)) ** 2).sum(0).mean()
    return cost, logp, al


def cost_and_grads(params, options, batch, alpha_c=0., decay_c=0., clip_c=0., dtype=torch.float64, **dp):
    """params: the init_params dict (numpy); batch: prepare_data's 8-tuple.  Returns
    (cost float, grads OrderedDict of numpy arrays in init_params order, extras dict)."""
    x, mask, ctxg, mask_ctxg, ctxl, _mask_ctxl, ctxm, _mask_ctxm = batch
    P = OrderedDict((k, _t(v, dtype).clone().requires_grad_(True)) for k, v in params.items())
    xt = torch.as_tensor(np.asarray(x), dtype=torch.long)
    args = [_t(a, dtype) for a in (mask, ctxg, mask_ctxg, ctxl, ctxm)]
    dpt = {k: _t(v, dtype) for k, v in dp.items() if v is not None}
    cost, logp, al = training_cost(P, options, xt, args[0], args[1], args[2], args[3], args[4], alpha_c, decay_c, **dpt)
    gl = torch.autograd.grad(cost, list(P.values()), allow_unused=True)
    grads = OrderedDict((k, (torch.zeros_like(P[k]) if g is None else g)) for k, g in zip(P.keys(), gl))
    g2 = sum(float((g ** 2).sum()) for g in grads.values())
    if clip_c > 0. and g2 > clip_c ** 2:                                            # :1194-1203
        grads = OrderedDict((k, g / np.sqrt(g2) * clip_c) for k, g in grads.items())
    return (float(cost.detach()), OrderedDict((k, g.detach().numpy()) for k, g in grads.items()),
            dict(logp=logp.detach().numpy(), g2=g2, **{k: v.detach().numpy() for k, v in al.items()}))
