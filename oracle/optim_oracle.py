"""CPU restatement (test infrastructure only) of the reference's parameter update: gradient clipping
(model_attention.py:1194-1203) and the optimizers of common.py:178-230, in float32 like the Theano graph
(floatX = float32).  Flat arrays stand for the per-tensor lists: the update is elementwise, and the clip
norm is a sum over all tensors.  Pinned to the reference: tests/golden/optim_ref.npz (trajectories from the
reference's own adam / adadelta source run on oracle/mini_theano) is replayed in tests/test_optim_golden.py."""
import numpy as np

f32 = np.float32


def clip(grads, clip_c):
    """g2 = sum g^2 over every tensor; switch(g2 > clip_c^2, g / sqrt(g2) * clip_c, g)   (:1195-1203)"""
    g = np.asarray(grads, f32)
    if clip_c <= 0:
        return g.copy(), f32((g.astype(np.float64) ** 2).sum())
    g2 = f32((g.astype(np.float64) ** 2).sum())          # the sum itself in high precision, then fp32 like the graph
    if g2 > f32(clip_c) ** 2:
        return (g / np.sqrt(g2) * f32(clip_c)).astype(f32), g2
    return g.copy(), g2


class Adam(object):
    """common.py:197-230.  lr0 = 0.0002 is hard-coded (the lr argument of f_update is unused); b1 = 0.1 and
    b2 = 0.001 multiply the new gradient; i counts updates from 0."""

    def __init__(self, n):
        self.m = np.zeros(n, f32)
        self.v = np.zeros(n, f32)
        self.i = f32(0.)

    def update(self, p, g):
        lr0, b1, b2, e = f32(0.0002), f32(0.1), f32(0.001), f32(1e-8)
        i_t = self.i + f32(1.)
        fix1 = f32(1.) - b1 ** i_t
        fix2 = f32(1.) - b2 ** i_t
        lr_t = lr0 * (np.sqrt(fix2) / fix1)
        g = np.asarray(g, f32)
        m_t = (b1 * g) + ((f32(1.) - b1) * self.m)
        v_t = (b2 * np.square(g)) + ((f32(1.) - b2) * self.v)
        g_t = m_t / (np.sqrt(v_t) + e)
        p_t = np.asarray(p, f32) - (lr_t * g_t)
        self.m, self.v, self.i = m_t.astype(f32), v_t.astype(f32), i_t
        return p_t.astype(f32)


class Adadelta(object):
    """common.py:178-195: rg2 is updated together with the gradient (f_grad_shared), the rest in f_update."""

    def __init__(self, n):
        self.rg2 = np.zeros(n, f32)
        self.ru2 = np.zeros(n, f32)

    def grad_shared(self, g):
        g = np.asarray(g, f32)
        self.rg2 = (f32(0.95) * self.rg2 + f32(0.05) * (g ** 2)).astype(f32)

    def update(self, p, g):
        g = np.asarray(g, f32)
        ud = -np.sqrt(self.ru2 + f32(1e-6)) / np.sqrt(self.rg2 + f32(1e-6)) * g
        self.ru2 = (f32(0.95) * self.ru2 + f32(0.05) * (ud ** 2)).astype(f32)
        return (np.asarray(p, f32) + ud).astype(f32)
