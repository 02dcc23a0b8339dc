"""Micro-benchmark of individual GEMM shapes through the kernel-level C-ABI (CUDA events, L2-exceeding operands)."""
import math, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch
from e2_tts_pytorch import _lib
from gpu_util import gemm, DEV

M = 100096
def run(name, N, K, epi, reps=5, **kw):
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    extra = {}
    if epi == _lib.EPI_RESID:
        out = torch.randn(M, N, device=DEV)
        extra = dict(out=out, ldo=N, resid=out, ldr=N, out_b16=torch.empty(M, N, device=DEV, dtype=torch.bfloat16) if kw.get('b16') else 0,
                     ldo_b16=N, gate=torch.rand(N, device=DEV), gate_bstride=0, lens=torch.full((128,), 782, device=DEV, dtype=torch.int32),
                     rows_per_batch=782)
    elif epi == _lib.EPI_F32:
        extra = dict(out=torch.empty(M, N, device=DEV), ldo=N)
    elif epi == _lib.EPI_BF16:
        extra = dict(out=torch.empty(M, N, device=DEV, dtype=torch.bfloat16), ldo=N)
    gemm(M, N, K, [a], w, epi, **extra)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        gemm(M, N, K, [a], w, epi, **extra)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f'{name:28s} N={N:5d} K={K:5d}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TF/s')

run('bf16 store', 1280, 1024, _lib.EPI_BF16)
run('f32 store', 1280, 1024, _lib.EPI_F32)
run('resid (f32 rmw)', 1280, 1024, _lib.EPI_RESID)
run('resid + bf16 copy', 1280, 1024, _lib.EPI_RESID, b16=True)
run('resid K=5120 + copy', 1280, 5120, _lib.EPI_RESID, b16=True)
run('resid N=1024 K=1024', 1024, 1024, _lib.EPI_RESID)
run('bf16 store N=3072', 3072, 1024, _lib.EPI_BF16)
