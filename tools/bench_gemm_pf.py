"""A/B of the A-operand L2 prefetch cursor (e2b_gemm_prefetch_kb) on the model's GEMM shapes at the C2 row count, kernel-level
C-ABI, CUDA events.  Between timed launches a 1 GB buffer is rewritten so that A is cold in L2, as it is inside the forward
(the previous kernel's output is 0.2-1 GB)."""
import ctypes as C, math, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch
from e2_tts_pytorch import _lib
from gpu_util import gemm, DEV
M = 100096
knob = C.c_int.in_dll(_lib.lib(), 'e2b_gemm_prefetch_kb')
rt_knob = C.c_int.in_dll(_lib.lib(), 'e2b_gemm_resid_tma')
pair_knob = C.c_int.in_dll(_lib.lib(), 'e2b_gemm_cta_pair')
MODE = os.environ.get('AB', 'resid')        # 'pf': A-operand prefetch distances; 'resid': TMA vs classic residual epilogue
flush = torch.empty(256 * 1024 * 1024, device=DEV)


def timeit(N, K, epi, extra, srcs):
    a = [torch.randn(M, k, device=DEV).to(torch.bfloat16) for k in srcs]
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    gemm(M, N, K, a, w, epi, **extra)
    tot = 0.0
    for _ in range(4):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        gemm(M, N, K, a, w, epi, **extra)
        e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / 4 * 1e3


def case(name, N, srcs, epi, mk):
    K = sum(srcs)
    row = f'{name:22s} N={N:5d} K={K:5d}:'
    if MODE == 'pf':
        for pf in (0, 6, 12, 24):
            knob.value = pf
            us = timeit(N, K, epi, mk(), srcs)
            row += f'  pf={pf:2d} {us:7.1f} us ({2 * M * N * K / us / 1e6:5.0f} TF/s)'
        knob.value = 0
    elif MODE == 'pair':
        ex = mk()
        for pr in (0, 1):
            pair_knob.value = pr
            us = timeit(N, K, epi, ex, srcs)
            var = (C.c_int * 4).in_dll(_lib.lib(), 'e2b_gemm_last_variant')
            row += f'  {"CTA pair" if pr else "1 CTA   "} {us:7.1f} us ({2 * M * N * K / us / 1e6:5.0f} TF/s) [epi {var[0]} bn {var[1]} ew {var[2]} cg {var[3]}]'
        pair_knob.value = 0
    else:
        ex = mk()
        nbytes = 2.0 * M * K + (8.0 * M * N if epi == _lib.EPI_RESID else 0) + (2.0 * M * N if 'out_b16' in ex else 0)
        for rt in ((0, 1) if epi == _lib.EPI_RESID else (1,)):
            rt_knob.value = rt
            us = timeit(N, K, epi, ex, srcs)
            row += f'  {"tma    " if rt else "classic"} {us:7.1f} us ({2 * M * N * K / us / 1e6:5.0f} TF/s, {nbytes / us / 1e3:5.0f} GB/s)'
        rt_knob.value = 1
    print(row, flush=True)


lens = torch.full((128,), 782, device=DEV, dtype=torch.int32)
def resid(N, b16=True):
    out = torch.randn(M, N, device=DEV)
    extra = dict(out=out, ldo=N, resid=out, ldr=N, gate=torch.rand(N, device=DEV), gate_bstride=0, lens=lens, rows_per_batch=782)
    if b16:
        extra.update(out_b16=torch.empty(M, N, device=DEV, dtype=torch.bfloat16), ldo_b16=N)
    return lambda: extra
def geglu(N):
    extra = dict(out=torch.empty(M, N // 2, device=DEV, dtype=torch.bfloat16), ldo=N // 2, bias=torch.randn(N, device=DEV))
    return lambda: extra
def qkv(H):
    HD = H * 64
    extra = dict(out=torch.empty(M, 2 * HD, device=DEV, dtype=torch.bfloat16), ldo=2 * HD, q_end=HD, k_end=2 * HD, v_end=3 * HD, q_scale=0.125,
                 rope=torch.randn(782, 32, 2, device=DEV), pos_off=0, rows_per_batch=782, vt=torch.zeros(M, HD, device=DEV, dtype=torch.bfloat16), vt_ld=HD,
                 heads_v=H, hgate=torch.empty(M, H, device=DEV), hgate_ld=H, hgate_bias=torch.zeros(H, device=DEV), v_rowmajor=1)
    return lambda: extra

case('geglu text', 10240, [1280], _lib.EPI_GEGLU, geglu(10240))
case('geglu audio', 8192, [1024], _lib.EPI_GEGLU, geglu(8192))
case('geglu frames', 4096, [512], _lib.EPI_GEGLU, geglu(4096))
case('ff2 text', 1280, [5120], _lib.EPI_RESID, resid(1280))
case('ff2 audio', 1024, [4096], _lib.EPI_RESID, resid(1024))
case('ff2 frames', 512, [2048], _lib.EPI_RESID, resid(512))
case('cross tfa', 1024, [1024, 1280, 512], _lib.EPI_RESID, resid(1024))
case('cross at', 1280, [1024, 1280], _lib.EPI_RESID, resid(1280, False))
case('cross af', 512, [1024, 512], _lib.EPI_RESID, resid(512, False))
case('out text', 1280, [1024], _lib.EPI_RESID, resid(1280, False))
case('out audio', 1024, [1024], _lib.EPI_RESID, resid(1024, False))
case('out frames', 512, [512], _lib.EPI_RESID, resid(512, False))
case('qkv text', 3088, [1280], _lib.EPI_QKV, qkv(16))
case('qkv audio', 3088, [1024], _lib.EPI_QKV, qkv(16))
case('qkv frames', 1544, [512], _lib.EPI_QKV, qkv(8))

if MODE == 'pair':
    # the audio out-projection as the engine launches it: per-clip gate, bf16(x * gain) with the gain switching at a row, row sums
    N = 1024
    gate_b = torch.rand(128, N, device=DEV)
    def fused(per_clip_gate, b16, scale, split, ss):
        out = torch.randn(M, N, device=DEV)
        ex = dict(out=out, ldo=N, resid=out, ldr=N, lens=lens, rows_per_batch=782)
        ex.update(dict(gate=gate_b, gate_bstride=N) if per_clip_gate else dict(gate=torch.rand(N, device=DEV), gate_bstride=0))
        if b16: ex.update(out_b16=torch.empty(M, N, device=DEV, dtype=torch.bfloat16), ldo_b16=N)
        if scale: ex.update(b16_scale=torch.rand(N, device=DEV))
        if split: ex.update(b16_scale2=torch.rand(N, device=DEV), b16_split_row=M // 2)
        if ss: ex.update(row_ss=torch.empty(16, M, device=DEV), row_ss_ld=M)
        return lambda: ex
    case('out audio +clip gate', N, [1024], _lib.EPI_RESID, fused(1, 0, 0, 0, 0))
    case('out audio +bf16', N, [1024], _lib.EPI_RESID, fused(1, 1, 0, 0, 0))
    case('out audio +gain', N, [1024], _lib.EPI_RESID, fused(1, 1, 1, 0, 0))
    case('out audio +split', N, [1024], _lib.EPI_RESID, fused(1, 1, 1, 1, 0))
    case('out audio +row sums', N, [1024], _lib.EPI_RESID, fused(1, 1, 1, 1, 1))
    case('out audio sums only', N, [1024], _lib.EPI_RESID, fused(1, 1, 0, 0, 1))
