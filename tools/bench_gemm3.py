"""A/B of the 8-epilogue-warp GEMM configuration on the model's shapes (kernel-level C-ABI, CUDA events)."""
import ctypes as C, math, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch
from e2_tts_pytorch import _lib
from gpu_util import gemm, DEV
M = 100096
knob = C.c_int.in_dll(_lib.lib(), 'e2b_gemm_ew8_max_k')

def timeit(N, K, epi, extra):
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    gemm(M, N, K, [a], w, epi, **extra)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): gemm(M, N, K, [a], w, epi, **extra)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 * 1e3

def case(name, N, K, epi, mk):
    res = []
    for thr in (0, 1 << 20):
        knob.value = thr
        us = timeit(N, K, epi, mk())
        res.append(us)
    print(f'{name:26s} N={N:5d} K={K:5d}: 4 warps/4 stages {res[0]:8.1f} us ({2*M*N*K/res[0]/1e6:6.0f} TF/s) | 8 warps/3 stages {res[1]:8.1f} us ({2*M*N*K/res[1]/1e6:6.0f} TF/s)')

lens = torch.full((128,), 782, device=DEV, dtype=torch.int32)
def resid(N):
    out = torch.randn(M, N, device=DEV)
    return lambda: dict(out=out, ldo=N, resid=out, ldr=N, out_b16=torch.empty(M, N, device=DEV, dtype=torch.bfloat16), ldo_b16=N,
                        gate=torch.rand(N, device=DEV), gate_bstride=0, lens=lens, rows_per_batch=782)
def geglu(N):
    return lambda: dict(out=torch.empty(M, N // 2, device=DEV, dtype=torch.bfloat16), ldo=N // 2, bias=torch.randn(N, device=DEV))
def qkv(H):
    HD = H * 64
    rope = torch.randn(782, 32, 2, device=DEV)
    return lambda: dict(out=torch.empty(M, 2 * HD, device=DEV, dtype=torch.bfloat16), ldo=2 * HD, q_end=HD, k_end=2 * HD, v_end=3 * HD, q_scale=0.125,
                        rope=rope, pos_off=0, rows_per_batch=782, vt=torch.zeros(128 * H * 64, 784, device=DEV, dtype=torch.bfloat16), vt_ld=784,
                        heads_v=H, hgate=torch.empty(M, H, device=DEV), hgate_ld=H, hgate_bias=torch.zeros(H, device=DEV))
case('geglu frames', 4096, 512, _lib.EPI_GEGLU, geglu(4096))
case('geglu audio', 8192, 1024, _lib.EPI_GEGLU, geglu(8192))
case('geglu text', 10240, 1280, _lib.EPI_GEGLU, geglu(10240))
case('resid frames out', 512, 512, _lib.EPI_RESID, resid(512))
case('resid audio out', 1024, 1024, _lib.EPI_RESID, resid(1024))
case('resid text out', 1280, 1024, _lib.EPI_RESID, resid(1280))
case('resid frames ff2', 512, 2048, _lib.EPI_RESID, resid(512))
case('resid text ff2', 1280, 5120, _lib.EPI_RESID, resid(1280))
case('qkv frames', 1544, 512, _lib.EPI_QKV, qkv(8))
case('qkv audio', 3088, 1024, _lib.EPI_QKV, qkv(16))
case('qkv text', 3088, 1280, _lib.EPI_QKV, qkv(16))
