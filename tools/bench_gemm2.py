import math, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch
from e2_tts_pytorch import _lib
from gpu_util import gemm, DEV
M, N, K = 100096, 1280, 1024
a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
def t(name, epi, **extra):
    gemm(M, N, K, [a], w, epi, **extra)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): gemm(M, N, K, [a], w, epi, **extra)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    print(f'{name:44s}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TF/s')
out = torch.randn(M, N, device=DEV); res = torch.randn(M, N, device=DEV)
lens = torch.full((128,), 782, device=DEV, dtype=torch.int32); gate = torch.rand(N, device=DEV)
t('f32 store', _lib.EPI_F32, out=out, ldo=N)
t('resid in-place, gate+lens', _lib.EPI_RESID, out=out, ldo=N, resid=out, ldr=N, gate=gate, gate_bstride=0, lens=lens, rows_per_batch=782)
t('resid in-place, no gate/lens', _lib.EPI_RESID, out=out, ldo=N, resid=out, ldr=N)
t('resid out-of-place, no gate/lens', _lib.EPI_RESID, out=out, ldo=N, resid=res, ldr=N)
small = torch.randn(782, N, device=DEV)
t('f32 + add_table (L2-resident per-row loads)', _lib.EPI_F32, out=out, ldo=N, rpb_in=782, rpb_out=782, row_off=0, add_table=small, ld_add=N)
tiny = torch.randn(1, N, device=DEV)
t('resid from a 1-row tensor (ldr=0, L1 hits)', _lib.EPI_RESID, out=out, ldo=N, resid=tiny, ldr=0)
