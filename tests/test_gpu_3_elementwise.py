"""Warp-level / element-wise kernels (csrc/elementwise.cu, csrc/melspec.cu) against torch."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import DEV, L, kcheck, rel
from e2_tts_pytorch import _lib
from e2_tts_pytorch.e2_tts_crossatt3 import MelSpec

sp = _lib.stream_ptr
P = _lib.ptr


@pytest.mark.parametrize('C_', [64, 192, 512, 1024, 1280, 2048])
def test_rmsnorm(C_):
    B, N, skip = 3, 37, 5
    x = torch.randn(B * N, C_, device=DEV) * 3
    x[4] = 0
    scale = torch.randn(B, C_, device=DEV)
    y = torch.zeros(B * (N - skip), C_, device=DEV, dtype=torch.bfloat16)
    kcheck(L().e2b_rmsnorm_launch(P(x), C_, P(y), C_, P(scale), C_, B, N, skip, C_, 0, sp()))
    ref = (F.normalize(x, dim=-1) * math.sqrt(C_)).reshape(B, N, C_)[:, skip:] * scale[:, None, :]
    assert rel(y.reshape(B, N - skip, C_), ref) < 4e-3
    yf = torch.zeros(B * N, C_, device=DEV)
    kcheck(L().e2b_rmsnorm_launch(P(x), C_, P(yf), C_, P(scale[0].contiguous()), 0, B, N, 0, C_, 1, sp()))
    ref2 = F.normalize(x, dim=-1) * math.sqrt(C_) * scale[0]
    assert rel(yf, ref2) < 1e-6 and torch.count_nonzero(yf[4]) == 0


@pytest.mark.parametrize('B,N,C_,lens', [(2, 82, 64, [82, 50]), (3, 782, 192, [782, 400, 33]), (1, 200, 1024, [200])])
def test_dwconv(B, N, C_, lens):
    x = torch.randn(B, N, C_, device=DEV)
    w = torch.randn(C_, 1, 31, device=DEV) / 5
    b = torch.randn(C_, device=DEV)
    y = torch.zeros_like(x)
    lt = torch.tensor(lens, device=DEV, dtype=torch.int32)
    wt = w[:, 0, :].t().contiguous()
    kcheck(L().e2b_dwconv_launch(P(x), P(y), P(wt), P(b), P(lt), B, N, C_, 31, sp()))
    mask = (torch.arange(N, device=DEV)[None, :] < lt[:, None])[..., None]
    xm = torch.where(mask, x, torch.zeros_like(x))
    c = F.silu(F.conv1d(xm.transpose(1, 2), w, b, padding=15, groups=C_)).transpose(1, 2)
    ref = x + torch.where(mask, c, torch.zeros_like(c))
    assert rel(y, ref) < 1e-5


@pytest.mark.parametrize('B,N,C_,lens', [(2, 100, 192, [100, 63]), (3, 782, 1280, [782, 500, 40]), (1, 65, 64, [65]), (2, 130, 512, [130, 129])])
def test_dwconv_with_normed_operand_outputs(B, N, C_, lens):
    """Norm as a row scale, producing side in the conv kernel: same y as the plain kernel (bit for bit), plus bf16(y * gain) and the
    sums of y^2 per row and 128-channel group."""
    x = torch.randn(B, N, C_, device=DEV)
    w = torch.randn(C_, 1, 31, device=DEV) / 5
    b = torch.randn(C_, device=DEV)
    gain = torch.rand(C_, device=DEV) + 0.5
    lt = torch.tensor(lens, device=DEV, dtype=torch.int32)
    wt = w[:, 0, :].t().contiguous()
    y0 = torch.zeros_like(x)
    kcheck(L().e2b_dwconv_launch(P(x), P(y0), P(wt), P(b), P(lt), B, N, C_, 31, sp()))
    raw, y, pad = _guarded((B, N, C_))
    rawb, yb, padb = _guarded((B * N, C_), torch.bfloat16)
    parts = (C_ + 127) // 128
    ss = torch.full((parts, B * N), float('nan'), device=DEV)
    kcheck(L().e2b_dwconv_norm_launch(P(x), P(y), P(wt), P(b), P(lt), B, N, C_, 31, P(yb), P(gain), P(ss), B * N, sp()))
    assert torch.equal(y, y0) and _guards_intact(raw, pad) and _guards_intact(rawb, padb)
    yf = y.reshape(B * N, C_)
    assert rel(yb, yf * gain) < 4e-3
    want = torch.stack([(yf[:, p * 128:(p + 1) * 128] ** 2).sum(1) for p in range(parts)])
    assert rel(ss, want) < 1e-5


def test_time_conditioning():
    dim, nt, nmat = 128, 5, 4
    g = torch.Generator().manual_seed(0)
    times = torch.rand(nt, generator=g).to(DEV)
    fw = torch.randn(dim // 2, generator=g).to(DEV)
    w1 = (torch.randn(dim, dim + 1, generator=g) / 10).to(DEV)
    b1 = torch.randn(dim, generator=g).to(DEV)
    tcond = torch.zeros(nt, dim, device=DEV)
    kcheck(L().e2b_time_mlp_launch(P(times), nt, P(fw), P(w1), P(b1), dim, P(tcond), sp()))
    fr = times[:, None] * fw[None, :] * 2 * math.pi
    ref = F.silu(F.linear(torch.cat((times[:, None], fr.sin(), fr.cos()), -1), w1, b1))
    assert rel(tcond, ref) < 1e-5
    ws = [(torch.randn(dim, dim, generator=g) / 10).to(DEV) for _ in range(nmat)]
    bs = [torch.randn(dim, generator=g).to(DEV) if m % 2 else None for m in range(nmat)]
    wp = torch.tensor([w.data_ptr() for w in ws], dtype=torch.int64, device=DEV)
    bp = torch.tensor([0 if b is None else b.data_ptr() for b in bs], dtype=torch.int64, device=DEV)
    act = torch.tensor([m % 2 for m in range(nmat)], dtype=torch.int32, device=DEV)
    out = torch.zeros(nt, nmat, dim, device=DEV)
    kcheck(L().e2b_time_gemv_launch(P(ref), nt, dim, P(wp), P(bp), P(act), nmat, P(out), sp()))
    for m in range(nmat):
        r = F.linear(ref, ws[m], bs[m])
        r = torch.sigmoid(r) if m % 2 else r + 1
        assert rel(out[:, m], r) < 1e-5


def test_init_stream_and_cast():
    B, Bt, n, R, C_ = 2, 4, 10, 32, 64
    regs = torch.randn(R, C_, device=DEV)
    src = torch.randn(B, n, C_, device=DEV)
    table = torch.randn(n, C_, device=DEV)
    drop = torch.tensor([0, 0, 1, 1], dtype=torch.uint8, device=DEV)
    dst = torch.full((Bt, R + n, C_), float('nan'), device=DEV)
    dstb = torch.zeros(Bt, R + n, C_, device=DEV, dtype=torch.bfloat16)
    kcheck(L().e2b_init_stream_launch(P(dst), P(dstb), P(regs), P(src), B, P(drop), P(table), Bt, n, R, C_, sp()))
    ref = torch.cat((regs.expand(Bt, -1, -1), torch.cat((src + table, table.expand(B, -1, -1)), 0)), 1)
    assert torch.equal(dst, ref) and rel(dstb, ref) < 4e-3
    dst2 = torch.full((Bt, R + n, C_), 7.0, device=DEV)
    kcheck(L().e2b_init_stream_launch(P(dst2), None, P(regs), None, -1, None, None, Bt, n, R, C_, sp()))
    assert torch.equal(dst2[:, :R], regs.expand(Bt, -1, -1)) and torch.all(dst2[:, R:] == 7.0)
    s = torch.randn(9, 51, device=DEV)
    d = torch.full((9, 64), float('nan'), device=DEV, dtype=torch.bfloat16)
    kcheck(L().e2b_cast_pad_launch(P(s), 51, P(d), 64, 9, 51, sp()))
    assert torch.equal(d[:, :51], s.to(torch.bfloat16)) and torch.count_nonzero(d[:, 51:]) == 0


def _project(x, y):
    sh = x.shape
    x, y = x.flatten(1).double(), y.flatten(1).double()
    unit = F.normalize(y, dim=-1)
    par = (x * unit).sum(-1, keepdim=True) * unit
    return par.reshape(sh).float(), (x - par).reshape(sh).float()


def test_guided_euler_cfg_kpass_and_apg():
    B, n, d = 3, 50, 64
    g = torch.Generator().manual_seed(1)
    y0 = torch.randn(B, n, d, generator=g).to(DEV)
    pred = torch.randn(4, B, n, d, generator=g).to(DEV)
    w = [2.0, 0.5, -0.25]
    y = y0.clone()
    yb = torch.zeros(2, B, n, d, device=DEV, dtype=torch.bfloat16)
    kcheck(L().e2b_guided_euler_launch(P(y), P(pred), 4, B, n * d, _lib.float_array(w), 0.125, 0, 0.0, None, P(yb), 2, sp()))
    v = pred[0] + sum(wk * (pred[0] - pred[k + 1]) for k, wk in enumerate(w))
    ref = y0 + 0.125 * v
    assert rel(y, ref) < 1e-6 and rel(yb[1], ref) < 4e-3
    # APG (remove_parallel_component): e2_tts_crossatt3.py:2106-2113
    y = y0.clone()
    scratch = torch.zeros(2 * B, device=DEV, dtype=torch.float64)
    kcheck(L().e2b_guided_euler_launch(P(y), P(pred), 2, B, n * d, _lib.float_array([2.0]), 0.1, 1, 0.25, P(scratch), None, 0, sp()))
    par, orth = _project(pred[0] - pred[1], pred[0])
    ref = y0 + 0.1 * (pred[0] + (orth + par * 0.25) * 2.0)
    assert rel(y, ref) < 1e-6
    # public wrapper
    y = y0.clone()
    kcheck(L().e2b_guided_euler(P(y), P(pred), 2, B, n * d, _lib.float_array([2.0]), 0.1, 0, 0.0, None, sp()))
    assert rel(y, y0 + 0.1 * (pred[0] + 2.0 * (pred[0] - pred[1]))) < 1e-6


def test_melspec_vs_torchaudio():
    import torchaudio
    from oracle import synth
    wav = torch.stack([synth.waveform(i, 24000 * 2 + 37) for i in range(3)]).to(DEV)
    ours = MelSpec()(wav)
    ref_mod = torchaudio.transforms.MelSpectrogram(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=100,
                                                   power=1, center=True, normalized=False, norm=None).to(DEV)
    ref = ref_mod(wav).clamp(min=1e-5).log()
    assert ours.shape == ref.shape
    assert (ours - ref).abs().max().item() < 2e-3 and rel(ours, ref) < 1e-4


def _guarded(shape, dtype=torch.float32, pad=4096):
    """A tensor carved out of a larger sentinel-filled buffer, to catch writes outside the output."""
    n = 1
    for s in shape:
        n *= s
    raw = torch.full((n + 2 * pad,), 12345.0, device=DEV, dtype=dtype)
    return raw, raw[pad:pad + n].view(*shape), pad


def _guards_intact(raw, pad):
    return bool((raw[:pad] == 12345.0).all() and (raw[-pad:] == 12345.0).all())


@pytest.mark.parametrize('B,N,C_', [(2, 83, 192), (3, 782, 64), (1, 65, 1280)])
def test_elementwise_kernels_stay_inside_their_outputs(B, N, C_):
    """No compute-sanitizer on this pool: every element-wise kernel writes into a guarded buffer instead."""
    lens = torch.tensor([N - 3 * b for b in range(B)], device=DEV, dtype=torch.int32)
    x = torch.randn(B, N, C_, device=DEV)
    # dwconv
    raw, y, pad = _guarded((B, N, C_))
    wt = (torch.randn(31, C_, device=DEV) / 5).contiguous()
    cb = torch.randn(C_, device=DEV)
    kcheck(L().e2b_dwconv_launch(P(x), P(y), P(wt), P(cb), P(lens), B, N, C_, 31, sp()))
    torch.cuda.synchronize()
    assert _guards_intact(raw, pad) and torch.isfinite(y).all()
    # rmsnorm, bf16 and fp32 outputs, with skipped leading rows
    scale = torch.rand(B, C_, device=DEV) + 0.5
    for mode, dt in ((0, torch.bfloat16), (1, torch.float32)):
        raw, yn, pad = _guarded((B * (N - 5), C_), dt)
        kcheck(L().e2b_rmsnorm_launch(P(x), C_, P(yn), C_, P(scale), C_, B, N, 5, C_, mode, sp()))
        torch.cuda.synchronize()
        assert _guards_intact(raw, pad) and torch.isfinite(yn.float()).all()
    # condition staging
    F_ = 17
    emb = torch.randn(B * F_, 1280, device=DEV)
    meta = torch.tensor([(b * F_, F_, max(1, N - 7 * b), 0) for b in range(B)], dtype=torch.int64, device=DEV)
    dur = torch.full((B,), 3.3, dtype=torch.float64, device=DEV)
    raw, out, pad = _guarded((B, N, 1280))
    kcheck(L().e2b_stage_clip(P(emb), P(meta), P(dur), B, N, 1280, 24000, 320, P(out), sp()))
    torch.cuda.synchronize()
    assert _guards_intact(raw, pad) and torch.isfinite(out).all()
