"""Time the depthwise-conv kernel at the C2 shapes (128 sequences x 782 rows) and check it against torch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch, torch.nn.functional as F
from gpu_util import L, kcheck, rel, DEV
from e2_tts_pytorch import _lib
sp, P = _lib.stream_ptr, _lib.ptr
B, N = 128, 782
for C_ in (1024, 1280, 512):
    x = torch.randn(B, N, C_, device=DEV)
    w = torch.randn(C_, 1, 31, device=DEV) / 5
    b = torch.randn(C_, device=DEV)
    y = torch.zeros_like(x)
    lt = torch.full((B,), N, device=DEV, dtype=torch.int32)
    lt[1] = 400
    wt = w[:, 0, :].t().contiguous()
    run = lambda: kcheck(L().e2b_dwconv_launch(P(x), P(y), P(wt), P(b), P(lt), B, N, C_, 31, sp()))
    run()
    mask = (torch.arange(N, device=DEV)[None, :] < lt[:, None])[..., None]
    xm = torch.where(mask, x, torch.zeros_like(x))
    ref = x + torch.where(mask, F.silu(F.conv1d(xm.transpose(1, 2), w, b, padding=15, groups=C_)).transpose(1, 2), torch.zeros_like(x))
    err = rel(y, ref)
    del ref, xm
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): run()
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    # the variant that also writes the normed bf16 operand and the row sums, against conv + rmsnorm as two kernels
    yb = torch.empty(B * N, C_, device=DEV, dtype=torch.bfloat16)
    gain = torch.rand(C_, device=DEV) + 0.5
    ss = torch.empty((C_ + 127) // 128, B * N, device=DEV)
    runn = lambda: kcheck(L().e2b_dwconv_norm_launch(P(x), P(y), P(wt), P(b), P(lt), B, N, C_, 31, P(yb), P(gain), P(ss), B * N, sp()))
    runr = lambda: kcheck(L().e2b_rmsnorm_launch(P(y), C_, P(yb), C_, P(gain), 0, B, N, 0, C_, 0, sp()))
    def tm(fn):
        for _ in range(3): fn()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10): fn()
        a1.record(); torch.cuda.synchronize()
        return a0.elapsed_time(a1) * 100
    usn, usr = tm(runn), tm(runr)
    print(f'   C={C_:5d}: conv+norm outputs {usn:7.1f} us  vs conv {us:7.1f} + rmsnorm {usr:7.1f} = {us + usr:7.1f} us')
    print(f'C={C_:5d}: {us:7.1f} us  {8.0 * B * N * C_ / us / 1e3:7.1f} GB/s  rel err {err:.1e}')
