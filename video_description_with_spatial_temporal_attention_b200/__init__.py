"""B200-native caption-decoding hot path of the spatial-temporal-attention video
captioner (reference: tuyunbin/Video-Description-with-Spatial-Temporal-Attention,
``model_attention.py``).  CUDA kernels + C ABI in ``csrc/`` (``libstat_b200.so``),
the reference-shaped host API in ``model_attention``.

Importing this package does not need a GPU; creating an ``Engine`` does.
"""
from . import common, synthetic  # noqa: F401
from ._lib import StatError, load as load_library  # noqa: F401


def default_options(**kw):
    """Option keys the hot path consumes (reference config.py:17-49 and the kwargs
    of train(), model_attention.py:1034-1078).  ``global_proj`` enables the
    ``ff_global`` layer the reference left commented out (:553-554, :661-662) so
    that ctxg_dim may differ from dim (decision D1)."""
    o = dict(dim_word=512, dim=512, ctxg_dim=512, ctxl_dim=4096, ctxm_dim=4096, ctxglm_dim=512,
             n_words=12594, selector=True, prev2out=True, ctx2out=True, use_dropout=True,
             n_layers_out=1, n_layers_init=0, encoder='none', global_proj=False)
    o.update(kw)
    if not o['global_proj'] and o['ctxg_dim'] != o['dim']:
        raise ValueError('the reference graph needs ctxg_dim == dim; set global_proj=True otherwise')
    o['ctxglm_dim'] = o['dim']
    return o


def baseline_options():
    """BASELINE.json dims: Dg=2048, Dm=Dr=4096, H=E=512, V=12594 (needs D1)."""
    return default_options(ctxg_dim=2048, global_proj=True)
