"""Helpers for the -m gpu tests: thin wrappers that fill the C-ABI descriptors from torch tensors."""
import ctypes as C

import torch

from e2_tts_pytorch import _lib

DEV = 'cuda:0'


def L():
    return _lib.lib()


def gemm(M, N, K, a_list, w, epi, **kw):
    d = _lib.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.num_src = len(a_list)
    for i, a in enumerate(a_list):
        d.a[i] = a.data_ptr(); d.lda[i] = a.stride(0); d.ka[i] = a.shape[1]
    d.w = w.data_ptr(); d.ldw = w.stride(0); d.epi = epi
    keep = []
    for k, v in kw.items():
        if torch.is_tensor(v):
            keep.append(v)
            setattr(d, k, v.data_ptr())
        else:
            setattr(d, k, v)
    rc = L().e2b_gemm_launch(C.byref(d), _lib.stream_ptr())
    if rc != 0:
        raise RuntimeError(L().e2b_kernel_last_error().decode())
    torch.cuda.synchronize()


def attention(**kw):
    d = _lib.AttnDesc()
    keep = []
    for k, v in kw.items():
        if torch.is_tensor(v):
            keep.append(v)
            setattr(d, k, v.data_ptr())
        else:
            setattr(d, k, v)
    rc = L().e2b_attention_launch(C.byref(d), _lib.stream_ptr())
    if rc != 0:
        raise RuntimeError(L().e2b_kernel_last_error().decode())
    torch.cuda.synchronize()


def kcheck(rc):
    if rc != 0:
        raise RuntimeError(L().e2b_kernel_last_error().decode())
    torch.cuda.synchronize()


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def bf(t):
    return t.to(torch.bfloat16)
