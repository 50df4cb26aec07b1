"""ctypes binding of ``libstat_b200.so`` (C ABI: ``include/stat_b200.h``).

There is no CPU path: if the shared library has not been built, or no CUDA
device is present, every compute entry point raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` (repo root) or
``python -m video_description_with_spatial_temporal_attention_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libstat_b200.so')

STAT_SELECTOR, STAT_PREV2OUT, STAT_CTX2OUT, STAT_GLOBAL_PROJ = 1, 2, 4, 8

# field order = include/stat_b200.h StatParams (the reference's init_params order
# with the D1 ff_global pair after ff_memory)
PARAM_FIELDS = (
    'Wemb', 'ff_state_W', 'ff_state_b', 'ff_memory_W', 'ff_memory_b',
    'ff_global_W', 'ff_global_b', 'ff_local_W', 'ff_local_b', 'ff_motion_W', 'ff_motion_b',
    'decoder_W', 'decoder_U', 'decoder_b', 'decoder_Wc',
    'decoder_Wcg_att', 'decoder_Wcm_att', 'decoder_Wclt_att',
    'decoder_Wdg_att', 'decoder_Wdm_att', 'decoder_Wdlt_att',
    'decoder_bg_att', 'decoder_bm_att', 'decoder_blt_att',
    'decoder_Wcl_att', 'decoder_Wdl_att', 'decoder_bl_att',
    'decoder_Ug_att', 'decoder_cg_att', 'decoder_Um_att', 'decoder_cm_att',
    'decoder_Ult_att', 'decoder_clt_att', 'decoder_Ul_att', 'decoder_cl_att',
    'decoder_W_sel', 'decoder_b_sel',
    'ff_logit_lstm_W', 'ff_logit_lstm_b', 'ff_logit_ctxglm_W', 'ff_logit_ctxglm_b',
    'ff_logit_W', 'ff_logit_b',
)


class StatDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('B', 'T', 'R', 'Dg', 'Dm', 'Dr', 'H', 'E', 'V', 'flags')]


class StatParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


class StatFwdBlocks(C.Structure):
    """include/stat_b200.h StatFwdBlocks: the forward's step-invariant blocks the backward pass reads"""
    FIELDS = ('ctxg0', 'pctxg', 'ctxm0', 'pctxm', 'ctxl0', 'pctxl', 'qctxl', 'h0c0')
    _fields_ = [(n, C.c_void_p) for n in FIELDS]


class StatError(RuntimeError):
    pass


_lib = None


def _declare(lib):
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    dp = C.POINTER(StatDims)
    lib.stat_version.restype = i32
    lib.stat_last_error.restype = C.c_char_p
    lib.stat_set_gemm_impl.argtypes = [i32]
    lib.stat_set_l2_persist.argtypes = [C.c_longlong]
    lib.stat_set_l2_persist.restype = C.c_longlong
    lib.stat_prepared_bytes.restype = sz
    lib.stat_prepared_bytes.argtypes = [dp]
    lib.stat_prepare_params.argtypes = [dp, C.POINTER(StatParams), vp, vp]
    lib.stat_workspace_bytes.restype = sz
    lib.stat_workspace_bytes.argtypes = [dp, i32]
    lib.stat_workspace_region.argtypes = [dp, i32, C.c_char_p, C.POINTER(sz), C.POINTER(sz)]
    lib.stat_precompute.argtypes = [dp, vp, vp, vp, vp, vp, vp, vp]
    lib.stat_init_state.argtypes = [dp] + [vp] * 7
    lib.stat_forward_teacher.argtypes = [dp, vp, vp, i32] + [vp] * 12
    lib.stat_decode_greedy.argtypes = [dp, vp, vp, i32, vp, vp, vp, vp]
    lib.stat_decode_beam.argtypes = [dp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.stat_step.argtypes = [dp, vp, vp, i32] + [vp] * 8
    lib.stat_attention.argtypes = [dp, vp, vp, i32, vp, vp]
    lib.stat_gemm.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, C.c_float, C.c_float, i32, i32, vp]
    lib.stat_launch_count.restype = C.c_ulonglong
    lib.stat_clip_scratch_bytes.restype = sz
    lib.stat_grad_clip.argtypes = [vp, sz, C.c_float, vp, vp, vp]
    lib.stat_adam_step.argtypes = [vp, vp, vp, vp, sz, i32, vp]
    lib.stat_alpha_coverage.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    lib.stat_adadelta_step.argtypes = [vp, vp, vp, vp, sz, i32, vp]
    lib.stat_grad_workspace_bytes.restype = sz
    lib.stat_grad_workspace_bytes.argtypes = [dp, i32]
    lib.stat_grad_shared.argtypes = ([dp, C.POINTER(StatParams), C.POINTER(StatFwdBlocks), i32] + [vp] * 14
                                     + [C.c_float] * 3 + [C.POINTER(StatParams), vp, vp])
    lib.stat_grad_profile_enable.argtypes = [i32]
    lib.stat_grad_profile_phases.restype = i32
    lib.stat_grad_profile_phase_name.restype = C.c_char_p
    lib.stat_grad_profile_phase_name.argtypes = [i32]
    lib.stat_grad_profile_collect.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int), i32]
    lib.stat_profile_enable.argtypes = [i32]
    lib.stat_profile_phases.restype = i32
    lib.stat_profile_phase_name.restype = C.c_char_p
    lib.stat_profile_phase_name.argtypes = [i32]
    lib.stat_profile_collect.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int), i32]
    lib.stat_set_step_impl.argtypes = [i32]
    lib.stat_set_beam_share.argtypes = [i32]
    lib.stat_set_beam_share.restype = i32
    for n in ('stat_set_step_impl', 'stat_attention', 'stat_profile_enable', 'stat_profile_collect', 'stat_set_gemm_impl', 'stat_prepare_params', 'stat_init_state', 'stat_workspace_region', 'stat_precompute',
              'stat_forward_teacher', 'stat_decode_greedy', 'stat_decode_beam', 'stat_step', 'stat_gemm', 'stat_grad_clip',
              'stat_adam_step', 'stat_adadelta_step', 'stat_alpha_coverage', 'stat_grad_shared', 'stat_grad_profile_enable', 'stat_grad_profile_collect'):
        getattr(lib, n).restype = i32


def load():
    """Load the shared library; raises StatError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise StatError('%s is missing: build it with __graft_entry__.build(); there is no CPU path'
                            % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        _declare(lib)
        impl = os.environ.get('STAT_GEMM_IMPL')          # debugging aid: 1 = fp32 SIMT GEMM kernel
        if impl is not None:
            lib.stat_set_gemm_impl(int(impl))
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().stat_last_error()
        raise StatError('libstat_b200 error %d: %s' % (rc, msg.decode() if msg else ''))


def flags_of(options):
    f = 0
    if options.get('selector'):
        f |= STAT_SELECTOR
    if options.get('prev2out'):
        f |= STAT_PREV2OUT
    if options.get('ctx2out'):
        f |= STAT_CTX2OUT
    if options.get('global_proj'):
        f |= STAT_GLOBAL_PROJ
    return f
