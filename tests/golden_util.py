"""Loader for the reference-generated fixtures in tests/golden/."""
import glob
import json
import os
from collections import OrderedDict

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'ref_*.npz')))
BATCH_KEYS = ('x', 'mask', 'ctxg', 'ctxg_mask', 'ctxl', 'ctxl_mask', 'ctxm', 'ctxm_mask')


class Golden(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.z = z
        self.options = json.loads(str(z['options']))
        order = json.loads(str(z['param_order']))
        self.params = OrderedDict((k, z['param:' + k]) for k in order)
        self.batch = tuple(z['in:' + k] for k in BATCH_KEYS)
        self.ks = [int(k) for k in z['gen_ks']]
        self.maxlen = int(z['gen_maxlen'])

    def out(self, k):
        return self.z['out:' + k]

    def inp(self, k):
        return self.z['in:' + k]

    def hyps(self, k, b):
        toks = self.z['out:gen_k%d_b%d_tokens' % (k, b)]
        return [[int(t) for t in row if t >= 0] for row in toks], \
            self.z['out:gen_k%d_b%d_scores' % (k, b)]
