"""ctypes binding of libe2b.so (include/e2b.h + csrc/kernels.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, 'libe2b.so')

DROP_CLIP, DROP_CTX, DROP_ROLL, DROP_AUDIO = 1, 2, 4, 8


class Config(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('depth', 'dim', 'dim_text', 'dim_frames', 'heads', 'dim_head', 'frames_heads',
                                       'num_channels', 'num_registers', 'kernel_size', 'notes', 'max_seq_len', 'ff_mult', 'precision')]


class Tensor(C.Structure):
    _fields_ = [('name', C.c_char_p), ('dev', C.c_void_p), ('ndim', C.c_int), ('shape', C.c_longlong * 4)]


class GemmDesc(C.Structure):
    _fields_ = [
        ('M', C.c_int), ('N', C.c_int), ('K', C.c_int), ('num_src', C.c_int),
        ('a', C.c_void_p * 9), ('lda', C.c_int * 9), ('ka', C.c_int * 9),
        ('w', C.c_void_p), ('ldw', C.c_int), ('epi', C.c_int),
        ('bias', C.c_void_p), ('out', C.c_void_p), ('ldo', C.c_int),
        ('out_b16', C.c_void_p), ('ldo_b16', C.c_int),
        ('resid', C.c_void_p), ('ldr', C.c_int),
        ('gate', C.c_void_p), ('gate_bstride', C.c_int),
        ('lens', C.c_void_p), ('rows_per_batch', C.c_int),
        ('rpb_in', C.c_int), ('rpb_out', C.c_int), ('row_off', C.c_int),
        ('add_table', C.c_void_p), ('ld_add', C.c_int),
        ('q_end', C.c_int), ('k_end', C.c_int), ('v_end', C.c_int), ('q_scale', C.c_float),
        ('rope', C.c_void_p), ('pos_off', C.c_int),
        ('vt', C.c_void_p), ('vt_ld', C.c_int), ('heads_v', C.c_int),
        ('hgate', C.c_void_p), ('hgate_ld', C.c_int), ('hgate_bias', C.c_void_p),
        ('split', C.c_int), ('qk_f32', C.c_void_p), ('v_f32', C.c_void_p), ('v_f32_ld', C.c_int), ('v_rowmajor', C.c_int),
        ('row_ss', C.c_void_p), ('row_ss_ld', C.c_int), ('b16_scale', C.c_void_p), ('b16_scale2', C.c_void_p), ('b16_split_row', C.c_int),
        ('in_row_ss', C.c_void_p), ('in_row_parts', C.c_int), ('in_row_ss_ld', C.c_int), ('in_row_mult', C.c_float),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ('batch', C.c_int), ('heads', C.c_int), ('q_rows_per_batch', C.c_int), ('kv_rows_per_batch', C.c_int),
        ('q', C.c_void_p), ('ldq', C.c_int), ('q_col0', C.c_int),
        ('k', C.c_void_p), ('ldk', C.c_int), ('k_col0', C.c_int),
        ('vt', C.c_void_p), ('vt_ld', C.c_int), ('kv_batch_mod', C.c_int),
        ('kv_lens', C.c_void_p), ('kv_lens_add', C.c_int),
        ('hgate', C.c_void_p), ('hgate_ld', C.c_int),
        ('out', C.c_void_p), ('ldo', C.c_int), ('softclamp', C.c_float), ('v_rowmajor', C.c_int), ('v_col0', C.c_int),
    ]


class AttnF32Desc(C.Structure):
    _fields_ = [
        ('batch', C.c_int), ('heads', C.c_int), ('q_rows_per_batch', C.c_int), ('kv_rows_per_batch', C.c_int),
        ('q', C.c_void_p), ('ldq', C.c_int), ('q_col0', C.c_int),
        ('k', C.c_void_p), ('ldk', C.c_int), ('k_col0', C.c_int),
        ('v', C.c_void_p), ('ldv', C.c_int), ('v_col0', C.c_int),
        ('kv_batch_mod', C.c_int), ('kv_lens', C.c_void_p), ('kv_lens_add', C.c_int),
        ('hgate', C.c_void_p), ('hgate_ld', C.c_int),
        ('out', C.c_void_p), ('ldo', C.c_int), ('out_split', C.c_int), ('softclamp', C.c_float),
    ]


EPI_BF16, EPI_F32, EPI_GEGLU, EPI_RESID, EPI_QKV = range(5)
PRECISIONS = {'bf16': 0, 'fp32': 1}

_SIGS = {
    # include/e2b.h
    'e2b_create': (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    'e2b_destroy': (None, [C.c_void_p]),
    'e2b_last_error': (C.c_char_p, [C.c_void_p]),
    'e2b_load_weights': (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.c_int, C.c_void_p]),
    'e2b_prepare': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    'e2b_set_conditions': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.c_void_p]),
    'e2b_set_audio_cond': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    'e2b_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    'e2b_sample': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float,
                             C.c_void_p]),
    'e2b_transformer_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]),
    'e2b_guided_euler': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.POINTER(C.c_float), C.c_float,
                                   C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    'e2b_melspec': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p]),
    'e2b_frame_windows': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
    'e2b_roll_expand': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'e2b_config_size': (C.c_int, []),
    'e2b_forward_flops': (C.c_double, [C.c_void_p]),
    'e2b_launch_count': (C.c_longlong, [C.c_void_p]),
    # kernel-level entry points (csrc/kernels.h) used by the unit tests
    'e2b_gemm_launch': (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    'e2b_gemm_row_parts': (C.c_int, [C.POINTER(GemmDesc)]),
    'e2b_attention_launch': (C.c_int, [C.POINTER(AttnDesc), C.c_void_p]),
    'e2b_attention_f32_launch': (C.c_int, [C.POINTER(AttnF32Desc), C.c_void_p]),
    'e2b_rmsnorm_launch': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p]),
    'e2b_conv1d_cl': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p]),
    'e2b_lstm_layer': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'e2b_stage_clip': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'e2b_dwconv_launch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p]),
    'e2b_dwconv_norm_launch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'e2b_time_mlp_launch': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'e2b_time_gemv_launch': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p]),
    'e2b_init_stream_launch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'e2b_cast_pad_launch': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'e2b_cast_part_launch': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'e2b_guided_euler_launch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.POINTER(C.c_float),
                                          C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'e2b_kernel_last_error': (C.c_char_p, []),
    'e2b_prof_enable': (None, [C.c_int]),
    'e2b_prof_report': (C.c_int, [C.c_char_p, C.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
_lib = None


def lib():
    """Load libe2b.so once.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found. Build it with `python {os.path.join(_PKG, "build.py")}` (needs nvcc, sm_100a). '
                'This package has no CPU / PyTorch-eager fallback.')
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.e2b_config_size() != C.sizeof(Config):
            raise RuntimeError(f'{LIB_PATH}: e2b_config is {L.e2b_config_size()} bytes in the library, {C.sizeof(Config)} in this binding '
                               '(stale build? run build.py --force)')
        _lib = L
    return _lib


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def check(rc, handle=None, what=''):
    if rc != 0:
        msg = lib().e2b_last_error(handle) if handle is not None else lib().e2b_last_error(None)
        kmsg = lib().e2b_kernel_last_error()
        raise RuntimeError(f'libe2b {what} failed: {(msg or b"").decode()} [{(kmsg or b"").decode()}]')


def int_array(vals):
    arr = (C.c_int * len(vals))(*[int(v) for v in vals])
    return arr


def float_array(vals):
    arr = (C.c_float * len(vals))(*[float(v) for v in vals])
    return arr


def profile_report():
    """Per-kernel CUDA-event timings recorded since e2b_prof_enable(1): list of dicts."""
    n = lib().e2b_prof_report(None, 0)
    buf = C.create_string_buffer(n + 16)
    lib().e2b_prof_report(buf, n + 16)
    rows = []
    for line in buf.value.decode().splitlines():
        kind, m, nn, k, cnt, ms, fl, by = line.split()
        rows.append(dict(kind=kind, m=int(m), n=int(nn), k=int(k), count=int(cnt), ms=float(ms), flops=float(fl), bytes=float(by)))
    return rows
