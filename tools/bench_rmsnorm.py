"""Time the rmsnorm kernel (fp32 rows -> bf16 normed rows) at the C2 shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch
from gpu_util import L, kcheck, DEV
from e2_tts_pytorch import _lib
sp, P = _lib.stream_ptr, _lib.ptr
B, N = 128, 782
for C_ in (1024, 1280, 512):
    x = torch.randn(B, N, C_, device=DEV)
    y = torch.empty(B, N, C_, device=DEV, dtype=torch.bfloat16)
    scale = torch.rand(B, C_, device=DEV) + 0.5
    run = lambda: kcheck(L().e2b_rmsnorm_launch(P(x), C_, P(y), C_, P(scale), C_, B, N, 0, C_, 0, sp()))
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    ref = torch.nn.functional.normalize(x, dim=-1) * (C_ ** 0.5) * scale[:, None, :]
    err = ((y.float() - ref).norm() / ref.norm()).item()
    print(f'C={C_:5d}: {us:7.1f} us  {6.0 * B * N * C_ / us / 1e3:7.1f} GB/s  rel err {err:.1e}')
