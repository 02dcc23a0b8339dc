"""tcgen05 GEMM (csrc/gemm.cu) against torch on the same bf16 operands."""
import ctypes
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import DEV, bf, gemm, rel
from e2_tts_pytorch import _lib


def _ab(M, N, K, seed=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    return a, w


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (128, 128, 64), (128, 256, 256), (300, 512, 1024), (1000, 192, 128),
                                   (77, 64, 192), (4096, 1280, 1024), (782, 3088, 1280)])
def test_plain_bf16(M, N, K):
    a, w = _ab(M, N, K)
    out = torch.full((M, N), float('nan'), device=DEV, dtype=torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    gemm(M, N, K, [a], w, _lib.EPI_BF16, out=out, ldo=N, bias=bias)
    ref = a.float() @ w.float().t() + bias
    e = rel(out, ref)
    print(f'gemm {M}x{N}x{K}: rel {e:.3e}')
    assert e < 5e-3


def test_diag_identity_patterns():
    """K-block / swizzle diagnostics: A = one-hot rows so each output column reads exactly one W element."""
    M, N, K = 128, 256, 128
    a = torch.zeros(M, K, device=DEV, dtype=torch.bfloat16)
    idx = torch.arange(M, device=DEV) % K
    a[torch.arange(M, device=DEV), idx] = 1
    w = bf(torch.arange(N * K, device=DEV, dtype=torch.float32).reshape(N, K) % 251 - 125)
    out = torch.zeros(M, N, device=DEV, dtype=torch.float32)
    gemm(M, N, K, [a], w, _lib.EPI_F32, out=out, ldo=N)
    ref = a.float() @ w.float().t()
    bad = (out != ref).nonzero()
    if bad.numel():
        print('first mismatches (row, col):', bad[:16].tolist())
        print('out[0,:8]', out[0, :8].tolist(), 'ref[0,:8]', ref[0, :8].tolist())
    assert torch.equal(out, ref)


def test_f32_remap_addtable_and_b16_copy():
    B, n, N_out, off, K, Nc = 3, 50, 82, 32, 64, 128
    a, w = _ab(B * n, Nc, K, 1)
    bias = torch.randn(Nc, device=DEV)
    table = torch.randn(n, Nc, device=DEV)
    out = torch.zeros(B * N_out, Nc, device=DEV)
    outb = torch.zeros(B * N_out, Nc, device=DEV, dtype=torch.bfloat16)
    gemm(B * n, Nc, K, [a], w, _lib.EPI_F32, out=out, ldo=Nc, out_b16=outb, ldo_b16=Nc, bias=bias, rpb_in=n, rpb_out=N_out,
         row_off=off, add_table=table, ld_add=Nc)
    ref = (a.float() @ w.float().t() + bias).reshape(B, n, Nc) + table
    got = out.reshape(B, N_out, Nc)
    assert rel(got[:, off:], ref) < 1e-5
    assert torch.count_nonzero(got[:, :off]) == 0
    assert rel(outb.reshape(B, N_out, Nc)[:, off:], ref) < 5e-3


def test_multi_source_concat_k():
    M = 500
    g = torch.Generator().manual_seed(2)
    a0, a1, a2 = (bf(torch.randn(M, k, generator=g)).to(DEV) for k in (128, 192, 64))
    w = bf(torch.randn(256, 384, generator=g) / 20).to(DEV)
    out = torch.zeros(M, 256, device=DEV)
    gemm(M, 256, 384, [a0, a1, a2], w, _lib.EPI_F32, out=out, ldo=256)
    ref = torch.cat([a0, a1, a2], 1).float() @ w.float().t()
    assert rel(out, ref) < 1e-5


def test_geglu_epilogue():
    M, dim, inner = 333, 128, 512
    g = torch.Generator().manual_seed(3)
    a = bf(torch.randn(M, dim, generator=g)).to(DEV)
    w = (torch.randn(2 * inner, dim, generator=g) / math.sqrt(dim)).to(DEV)
    b = torch.randn(2 * inner, generator=g).to(DEV)
    # pack: tile t = [value rows t*128.., gate rows inner + t*128..]
    order = torch.cat([torch.cat([torch.arange(t * 128, t * 128 + 128), inner + torch.arange(t * 128, t * 128 + 128)])
                       for t in range(inner // 128)]).to(DEV)
    wp, bp = bf(w[order]).contiguous(), b[order].contiguous()
    out = torch.zeros(M, inner, device=DEV, dtype=torch.bfloat16)
    gemm(M, 2 * inner, dim, [a], wp, _lib.EPI_GEGLU, out=out, ldo=inner, bias=bp)
    h = a.float() @ bf(w).float().t() + b
    ref = h[:, :inner] * torch.nn.functional.gelu(h[:, inner:])
    assert rel(out, ref) < 5e-3


@pytest.mark.parametrize('dim', [128, 1280])          # 8 epilogue warps (K <= 1536) in both; one and several K blocks
def test_geglu_and_qkv_row_scale(dim):
    """Norm as a row scale, consuming side: accumulator rows times sqrt(C) / max(||x||, 1e-12) with ||x||^2 given as partial sums,
    before bias / GELU (GEGLU) and before RoPE / q scale / V / head gate (QKV) -- equal to feeding the normalised rows."""
    M, inner, H, Nseq = 333, 512, 2, 111
    g = torch.Generator().manual_seed(dim)
    x = torch.randn(M, dim, generator=g).to(DEV) * (torch.rand(M, 1, generator=g).to(DEV) * 4 + 0.1)
    x[5] = 0                                                              # a zero row: the eps branch (scale sqrt(C) / 1e-12, result 0)
    a = bf(x)
    parts = 3
    ssq = (x * x).sum(1)
    split = torch.rand(parts, M, device=DEV)
    ss = (split / split.sum(0, keepdim=True) * ssq[None]).contiguous()
    rs = math.sqrt(dim) / ssq.sqrt().clamp_min(1e-12)
    an = a.float() * rs[:, None]                                          # what an rmsnorm with unit gain would have fed
    w = (torch.randn(2 * inner, dim, generator=g) / math.sqrt(dim)).to(DEV)
    b = torch.randn(2 * inner, generator=g).to(DEV)
    order = torch.cat([torch.cat([torch.arange(t * 128, t * 128 + 128), inner + torch.arange(t * 128, t * 128 + 128)])
                       for t in range(inner // 128)]).to(DEV)
    wp, bp = bf(w[order]).contiguous(), b[order].contiguous()
    out = torch.zeros(M, inner, device=DEV, dtype=torch.bfloat16)
    gemm(M, 2 * inner, dim, [a], wp, _lib.EPI_GEGLU, out=out, ldo=inner, bias=bp, in_row_ss=ss, in_row_parts=parts, in_row_ss_ld=M,
         in_row_mult=math.sqrt(dim))
    h = an @ bf(w).float().t() + b
    assert rel(out, h[:, :inner] * torch.nn.functional.gelu(h[:, inner:])) < 5e-3

    HD = H * 64
    wq = bf(torch.randn(3 * HD + H, dim, generator=g) / math.sqrt(dim)).to(DEV)
    hb = torch.randn(H, generator=g).to(DEV)
    ang = torch.arange(Nseq).float()[:, None] * (1. / (10000 ** (torch.arange(0, 64, 2).float() / 64)))[None, :]
    rope = torch.stack((ang.cos(), ang.sin()), -1).to(DEV).contiguous()
    for rows in (1, 0):
        qk = torch.zeros(M, 2 * HD, device=DEV, dtype=torch.bfloat16)
        hg = torch.zeros(M, H, device=DEV)
        vbuf = torch.zeros(M, HD, device=DEV, dtype=torch.bfloat16) if rows else torch.zeros(3 * H * 64, 112, device=DEV, dtype=torch.bfloat16)
        gemm(M, 3 * HD + H, dim, [a], wq, _lib.EPI_QKV, out=qk, ldo=2 * HD, q_end=HD, k_end=2 * HD, v_end=3 * HD, q_scale=0.125, rope=rope,
             pos_off=0, rows_per_batch=Nseq, vt=vbuf, vt_ld=HD if rows else 112, heads_v=H, hgate=hg, hgate_ld=H, hgate_bias=hb,
             v_rowmajor=rows, in_row_ss=ss, in_row_parts=parts, in_row_ss_ld=M, in_row_mult=math.sqrt(dim))
        full = an @ wq.float().t()

        def rot(t):
            t = t.reshape(3, Nseq, H, 32, 2)
            c, sn = rope[None, :, None, :, 0], rope[None, :, None, :, 1]
            return torch.stack((t[..., 0] * c - t[..., 1] * sn, t[..., 1] * c + t[..., 0] * sn), -1).reshape(M, HD)

        assert rel(qk[:, :HD], rot(full[:, :HD]) * 0.125) < 5e-3 and rel(qk[:, HD:], rot(full[:, HD:2 * HD])) < 5e-3
        v = full[:, 2 * HD:3 * HD]
        if rows:
            assert rel(vbuf, v) < 5e-3
        else:
            assert rel(vbuf[:, :Nseq], v.reshape(3, Nseq, H, 64).permute(0, 2, 3, 1).reshape(3 * H * 64, Nseq)) < 5e-3
        assert rel(hg, torch.sigmoid(full[:, 3 * HD:] + hb)) < 1e-4


def test_resid_gate_mask_epilogue():
    B, Nseq, K, C = 3, 70, 128, 192
    M = B * Nseq
    a, w = _ab(M, C, K, 4)
    bias = torch.randn(C, device=DEV)
    resid = torch.randn(M, C, device=DEV)
    gate = torch.rand(B, C, device=DEV)
    lens = torch.tensor([70, 41, 55], device=DEV, dtype=torch.int32)
    out = resid.clone()
    outb = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    gemm(M, C, K, [a], w, _lib.EPI_RESID, out=out, ldo=C, out_b16=outb, ldo_b16=C, resid=out, ldr=C, bias=bias, gate=gate,
         gate_bstride=C, lens=lens, rows_per_batch=Nseq)
    y = (a.float() @ w.float().t() + bias).reshape(B, Nseq, C) * gate[:, None, :]
    valid = (torch.arange(Nseq, device=DEV)[None, :] < lens[:, None])[..., None]
    ref = resid.reshape(B, Nseq, C) + torch.where(valid, y, torch.zeros_like(y))
    assert rel(out.reshape(B, Nseq, C), ref) < 1e-5
    assert rel(outb, ref.reshape(M, C)) < 5e-3


@pytest.mark.parametrize('tma', [1, 0], ids=['tma_epilogue', 'classic_epilogue'])
@pytest.mark.parametrize('B,Nseq,K,C,per_batch,b16,bias_on', [
    (30, 170, 128, 512, True, True, True),      # 2 column tiles, rows not a multiple of 128, per-batch gate rows
    (9, 300, 1024, 1024, False, True, False),   # 8 warps / two residual tiles (K <= 3072), shared gate vector, no bias
    (14, 150, 3200, 1280, False, True, True),   # 4 warps / one residual tile (K > 3072)
    (24, 90, 256, 1100, False, False, True),    # ragged last column tile (1100 = 4 x 256 + 76: a chunk of 12 live columns), no bf16 copy
    (3, 782, 512, 1056, True, True, False),     # last tile of one 32-column chunk: the second warp group has no chunk there
    (2, 100, 256, 512, True, True, True),       # few rows: 128-wide tiles (one wave), 4 warps / one residual tile
    (1, 77, 64, 192, False, True, False),       # narrow N, a 64-column last tile
])
def test_resid_epilogue_wide_tiles(B, Nseq, K, C, per_batch, b16, bias_on, tma):
    """EPI_RESID on 256-wide tiles, both epilogues (E2B_RESID_TMA): the residual stream moved by the TMA unit in a row-per-lane
    layout, and the classic load / transpose / store one.  In place (out aliases resid), row mask, gate, bias, bf16 copy."""
    import ctypes as C_
    knob = C_.c_int.in_dll(_lib.lib(), 'e2b_gemm_resid_tma')
    M = B * Nseq
    a, w = _ab(M, C, K, B + Nseq)
    bias = torch.randn(C, device=DEV) if bias_on else None
    resid = torch.randn(M, C, device=DEV)
    gate = torch.rand(B if per_batch else 1, C, device=DEV)
    lens = torch.tensor([Nseq - 3 * (i % 20) for i in range(B)], device=DEV, dtype=torch.int32)
    ss = torch.full(((C + 127) // 128, M), float('nan'), device=DEV) if tma else None
    raw = torch.full((M + 2, C), 777.0, device=DEV)                      # guard rows around the in-place stream
    out = raw[1:M + 1]
    out.copy_(resid)
    outb = torch.full((M + 1, C), -5.0, device=DEV, dtype=torch.bfloat16)
    kw = dict(out=out, ldo=C, resid=out, ldr=C, gate=gate, gate_bstride=C if per_batch else 0, lens=lens, rows_per_batch=Nseq)
    if bias_on:
        kw['bias'] = bias
    if b16:
        kw.update(out_b16=outb, ldo_b16=C)
    if tma:
        kw.update(row_ss=ss, row_ss_ld=M)
    knob.value = tma
    try:
        gemm(M, C, K, [a], w, _lib.EPI_RESID, **kw)
    finally:
        knob.value = 1
    variant = (C_.c_int * 3).in_dll(_lib.lib(), 'e2b_gemm_last_variant')
    wide = M * ((C + 127) // 128) > 128 * 148 and (C % 256 == 0 or C > 1024)
    assert list(variant)[:2] == [5 if tma else _lib.EPI_RESID, 256 if wide else 128], list(variant)   # the epilogue under test really ran
    assert variant[2] == ((8 if K <= (3072 if tma else 1536) else 4) if wide else 4)
    y = (a.float() @ w.float().t() + (bias if bias_on else 0)).reshape(B, Nseq, C) * (gate[:, None, :] if per_batch else gate[None])
    valid = (torch.arange(Nseq, device=DEV)[None, :] < lens[:, None])[..., None]
    ref = (resid.reshape(B, Nseq, C) + torch.where(valid, y, torch.zeros_like(y))).reshape(M, C)
    assert rel(out, ref) < 1e-5
    assert torch.equal(out[~valid.expand(B, Nseq, C).reshape(M, C)], resid[~valid.expand(B, Nseq, C).reshape(M, C)])   # masked rows untouched
    assert (raw[0] == 777.0).all() and (raw[M + 1] == 777.0).all()
    if b16:
        assert rel(outb[:M], ref) < 5e-3 and (outb[M] == -5.0).all()
    if tma:
        # one partial per 128 columns, whatever the tile configuration: partial p = sum of squares of columns [128 p, 128 p + 128)
        want = torch.stack([(out[:, p * 128:(p + 1) * 128] ** 2).sum(1) for p in range(ss.shape[0])])
        assert rel(ss, want) < 1e-5


@pytest.mark.parametrize('K,ew', [(1024, 8), (3200, 4)])
def test_cta_pair_variants_match_single_cta(K, ew):
    """tcgen05 cta_group::2 (a CTA pair per 256-row tile, each CTA staging half of the B tile): GEGLU and TMA-residual GEMMs must
    give bit-identical results to the one-CTA-per-tile kernels (same accumulation order), odd row-tile counts included."""
    knob = ctypes.c_int.in_dll(_lib.lib(), 'e2b_gemm_cta_pair')
    variant = (ctypes.c_int * 4).in_dll(_lib.lib(), 'e2b_gemm_last_variant')
    M, C, inner = 128 * 47 + 57, 1024, 2048                               # 48 row tiles of 128 -> 24 pairs; the last rows are ragged
    g = torch.Generator(device='cpu').manual_seed(K)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = bf(torch.randn(C, K, generator=g) / math.sqrt(K)).to(DEV)
    resid = torch.randn(M, C, generator=g).to(DEV)
    gate, bias = torch.rand(1, C, device=DEV), torch.randn(C, device=DEV)
    wg = bf(torch.randn(2 * inner, K, generator=g) / math.sqrt(K)).to(DEV)
    bg = torch.randn(2 * inner, device=DEV)
    H = 16
    HD = H * 64
    wq = bf(torch.randn(3 * HD + H, K, generator=g) / math.sqrt(K)).to(DEV)     # N = 3088: a ragged last tile of 16 columns
    hbq = torch.randn(H, device=DEV)
    rope = torch.randn(200, 32, 2, device=DEV)
    res = {}
    before = knob.value
    for pair in (0, 1):
        knob.value = pair
        try:
            out = resid.clone()
            outb = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
            ss = torch.zeros(C // 128, M, device=DEV)
            gemm(M, C, K, [a], w, _lib.EPI_RESID, out=out, ldo=C, resid=out, ldr=C, gate=gate, gate_bstride=0, bias=bias, out_b16=outb, ldo_b16=C,
                 row_ss=ss, row_ss_ld=M)
            v1 = list(variant)
            og = torch.zeros(M, inner, device=DEV, dtype=torch.bfloat16)
            gemm(M, 2 * inner, K, [a], wg, _lib.EPI_GEGLU, out=og, ldo=inner, bias=bg)
            v2 = list(variant)
            qk = torch.zeros(M, 2 * HD, device=DEV, dtype=torch.bfloat16)
            vr = torch.zeros(M, HD, device=DEV, dtype=torch.bfloat16)
            hg = torch.zeros(M, H, device=DEV)
            gemm(M, 3 * HD + H, K, [a], wq, _lib.EPI_QKV, out=qk, ldo=2 * HD, q_end=HD, k_end=2 * HD, v_end=3 * HD, q_scale=0.125, rope=rope,
                 pos_off=0, rows_per_batch=200, vt=vr, vt_ld=HD, heads_v=H, hgate=hg, hgate_ld=H, hgate_bias=hbq, v_rowmajor=1)
            v3 = list(variant)
            of = torch.zeros(M, C, device=DEV)
            gemm(M, C, K, [a], w, _lib.EPI_F32, out=of, ldo=C, bias=bias)
            v4 = list(variant)
        finally:
            knob.value = before
        assert v1 == [5, 256, 8 if K <= 3072 else 4, 2 if pair else 1] and v2 == [_lib.EPI_GEGLU, 256, 8 if K <= 1536 else 4, 2 if pair else 1], (v1, v2)
        assert v3[0] == 6 and v3[3] == (2 if pair else 1) and v4[0] == _lib.EPI_F32 and v4[3] == (2 if pair else 1), (v3, v4)
        res[pair] = (out, outb, ss, og, qk, vr, hg, of)
    ref = resid + (a.float() @ w.float().t() + bias) * gate
    assert rel(res[1][0], ref) < 1e-5
    assert rel(res[1][6], torch.sigmoid((a.float() @ wq.float().t())[:, 3 * HD:] + hbq)) < 1e-4      # the ragged tile's columns
    for x, y in zip(res[0], res[1]):
        assert torch.equal(x, y)


@pytest.mark.parametrize('B,H,K,cross,scaled', [(40, 16, 1024, 0, 1), (40, 8, 512, 0, 0), (3, 16, 1024, 0, 1), (40, 16, 1024, 1, 1), (17, 16, 2048, 0, 0), (17, 4, 1024, 0, 1)])
def test_qkv_tma_store_epilogue_matches_classic(B, H, K, cross, scaled):
    """The row-per-lane / TMA-store QKV epilogue (internal id 6: rope values of a lane's row kept in registers, bf16 tiles leave by
    TMA) against the classic transpose epilogue: bit-identical q | k, V rows and head gates -- for 256-wide tiles with 8 and 4
    epilogue warps, CTA pairs, the 128-wide tiles of a few-row launch, a q-only (cross-attention) projection, with and without the
    norm row scale; rows past M and the columns past N of the ragged tile must stay untouched."""
    knob = ctypes.c_int.in_dll(_lib.lib(), 'e2b_gemm_qkv_tma')
    variant = (ctypes.c_int * 4).in_dll(_lib.lib(), 'e2b_gemm_last_variant')
    Nseq = 203
    M, HD = B * Nseq, H * 64
    g = torch.Generator(device='cpu').manual_seed(B * 1000 + H)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    N = (HD if cross else 3 * HD) + H
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    hb = torch.randn(H, generator=g).to(DEV)
    ang = (torch.arange(Nseq + 5).float()[:, None] * (1. / (10000 ** (torch.arange(0, 64, 2).float() / 64)))[None, :])
    rope = torch.stack((ang.cos(), ang.sin()), -1).to(DEV).contiguous()
    ss = (torch.rand(3, M, generator=g) * K / 3).to(DEV)
    extra = dict(in_row_ss=ss, in_row_parts=3, in_row_ss_ld=M, in_row_mult=math.sqrt(K)) if scaled else {}
    qcols = HD if cross else 2 * HD
    res = {}
    before = knob.value
    for tma in (0, 1):
        knob.value = tma
        try:
            qk = torch.full((M + 3, qcols), 7.0, device=DEV, dtype=torch.bfloat16)
            vr = torch.full((M + 3, HD), 7.0, device=DEV, dtype=torch.bfloat16)
            hg = torch.full((M + 3, H), 7.0, device=DEV)
            gemm(M, N, K, [a], w, _lib.EPI_QKV, out=qk, ldo=qcols, q_end=HD, k_end=qcols, v_end=qcols if cross else 3 * HD, q_scale=0.125,
                 rope=rope, pos_off=5, rows_per_batch=Nseq, vt=vr, vt_ld=HD, heads_v=H, hgate=hg, hgate_ld=H, hgate_bias=hb, v_rowmajor=1, **extra)
            v = list(variant)
        finally:
            knob.value = before
        assert v[0] == (6 if tma else _lib.EPI_QKV), v
        if tma:
            bn = 128 if (B == 3 or H == 4) else 256                      # few rows / N = 772: 128-wide tiles
            assert v[1] == bn and v[2] == (4 if (K > 1536 or bn == 128) else 8), v
        res[tma] = (qk, vr, hg)
    for x, y in zip(res[0], res[1]):
        assert torch.equal(x, y)
    assert (res[1][0][M:] == 7.0).all() and (res[1][1][M:] == 7.0).all() and (res[1][2][M:] == 7.0).all()
    full = a.float() @ w.float().t()
    if scaled:
        full = full * (math.sqrt(K) / ss.sum(0).sqrt())[:, None]
    assert rel(res[1][2][:M], torch.sigmoid(full[:, N - H:] + hb)) < 1e-4
    if not cross:
        assert rel(res[1][1][:M], full[:, 2 * HD:3 * HD]) < 5e-3
    t = full[:, :qcols].reshape(B, Nseq, qcols // 64, 32, 2)
    c, sn = rope[None, 5:5 + Nseq, None, :, 0], rope[None, 5:5 + Nseq, None, :, 1]
    rot = torch.stack((t[..., 0] * c - t[..., 1] * sn, t[..., 1] * c + t[..., 0] * sn), -1).reshape(M, qcols)
    rot[:, :HD] *= 0.125
    assert rel(res[1][0][:M], rot) < 5e-3


def test_resid_row_sums_do_not_depend_on_the_tiling():
    """The same rows give bit-identical row-sum partials and results whether the launch takes 128-wide tiles (few rows) or 256-wide
    tiles with 8 or 4 epilogue warps (many rows): a clip's latent must not depend on the batch it is sampled in."""
    Nseq, K, C = 200, 1024, 1024
    outs = []
    for B in (2, 12):
        M = B * Nseq
        g = torch.Generator(device='cpu').manual_seed(77)
        a1 = bf(torch.randn(2 * Nseq, K, generator=g)).to(DEV)
        w = bf(torch.randn(C, K, generator=g) / math.sqrt(K)).to(DEV)
        r1 = torch.randn(2 * Nseq, C, generator=g).to(DEV)
        a = torch.cat([a1] + [torch.randn(Nseq, K, device=DEV).to(torch.bfloat16) for _ in range(B - 2)])
        out = torch.cat([r1] + [torch.randn(Nseq, C, device=DEV) for _ in range(B - 2)]).contiguous()
        ss = torch.zeros(8, M, device=DEV)
        outb = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
        gemm(M, C, K, [a], w, _lib.EPI_RESID, out=out, ldo=C, resid=out, ldr=C, out_b16=outb, ldo_b16=C, row_ss=ss, row_ss_ld=M)
        variant = list((ctypes.c_int * 3).in_dll(_lib.lib(), 'e2b_gemm_last_variant'))
        outs.append((variant, out[:2 * Nseq].clone(), ss[:, :2 * Nseq].clone(), outb[:2 * Nseq].clone()))
    assert outs[0][0][1] == 128 and outs[1][0][1] == 256, (outs[0][0], outs[1][0])
    for x, y in zip(outs[0][1:], outs[1][1:]):
        assert torch.equal(x, y)


def test_resid_epilogue_row_sums_and_scaled_copy():
    """Norm as a row scale: the TMA residual epilogue also emits sum(out^2) per row (partials per column tile and warp group) and
    bf16(out * scale[col]) with a second scale vector from a split row on."""
    B, Nseq, K, C = 9, 333, 1024, 1024
    M = B * Nseq
    a, w = _ab(M, C, K, 11)
    resid = torch.randn(M, C, device=DEV)
    gate = torch.rand(1, C, device=DEV)
    s1, s2 = torch.rand(C, device=DEV) + 0.5, torch.rand(C, device=DEV) + 0.5
    split = 1400
    out = resid.clone()
    outb = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    d = _lib.GemmDesc()
    d.N, d.K = C, K
    parts = _lib.lib().e2b_gemm_row_parts(ctypes.byref(d))
    assert parts == 8                                                     # 4 column tiles x 2 warp groups
    ss = torch.full((parts, M), float('nan'), device=DEV)
    gemm(M, C, K, [a], w, _lib.EPI_RESID, out=out, ldo=C, resid=out, ldr=C, gate=gate, gate_bstride=0, out_b16=outb, ldo_b16=C,
         row_ss=ss, row_ss_ld=M, b16_scale=s1, b16_scale2=s2, b16_split_row=split)
    ref = resid + (a.float() @ w.float().t()) * gate
    assert rel(out, ref) < 1e-5
    assert rel(ss.sum(0), (ref * ref).sum(1)) < 1e-5
    scale = torch.where(torch.arange(M, device=DEV)[:, None] >= split, s2[None], s1[None])
    assert rel(outb, ref * scale) < 5e-3


def test_qkv_epilogue_rope_vt_gate():
    B, Nseq, C, H = 2, 90, 128, 2
    HD, M = H * 64, 2 * 90
    g = torch.Generator().manual_seed(5)
    a = bf(torch.randn(M, C, generator=g)).to(DEV)
    w = bf(torch.randn(3 * HD + H, C, generator=g) / math.sqrt(C)).to(DEV)
    hb = torch.randn(H, generator=g).to(DEV)
    inv = 1. / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    pos = torch.arange(Nseq + 7).float()
    ang = pos[:, None] * inv[None, :]
    rope = torch.stack((ang.cos(), ang.sin()), -1).to(DEV).contiguous()          # [pos, 32, 2]
    Npad = 96
    qk = torch.zeros(M, 2 * HD, device=DEV, dtype=torch.bfloat16)
    vt = torch.zeros(B * H * 64, Npad, device=DEV, dtype=torch.bfloat16)
    hg = torch.zeros(M, H, device=DEV)
    gemm(M, 3 * HD + H, C, [a], w, _lib.EPI_QKV, out=qk, ldo=2 * HD, q_end=HD, k_end=2 * HD, v_end=3 * HD, q_scale=0.125,
         rope=rope, pos_off=7, rows_per_batch=Nseq, vt=vt, vt_ld=Npad, heads_v=H, hgate=hg, hgate_ld=H, hgate_bias=hb)
    full = a.float() @ w.float().t()
    q, k, v, gt = full[:, :HD], full[:, HD:2 * HD], full[:, 2 * HD:3 * HD], full[:, 3 * HD:]

    def rot(t):
        t = t.reshape(B, Nseq, H, 32, 2)
        c = rope[7:7 + Nseq, :, 0][None, :, None, :]
        s = rope[7:7 + Nseq, :, 1][None, :, None, :]
        x0, x1 = t[..., 0], t[..., 1]
        return torch.stack((x0 * c - x1 * s, x1 * c + x0 * s), -1).reshape(M, HD)

    assert rel(qk[:, :HD], rot(q) * 0.125) < 5e-3
    assert rel(qk[:, HD:], rot(k)) < 5e-3
    vref = v.reshape(B, Nseq, H, 64).permute(0, 2, 3, 1).reshape(B * H * 64, Nseq)
    assert rel(vt[:, :Nseq], vref) < 5e-3
    assert torch.count_nonzero(vt[:, Nseq:]) == 0
    assert rel(hg, torch.sigmoid(gt + hb)) < 1e-4
    # v as plain rows (the MN-major operand of the current attention kernel) + a wider N with a ragged last tile that only
    # holds the head-gate columns (the narrow-MMA path: N = 3 * 128 + 2 -> tiles of 256 | 130 valid columns)
    vrows = torch.full((M + 1, HD), -3.0, device=DEV, dtype=torch.bfloat16)
    qk2, hg2 = torch.zeros_like(qk), torch.zeros_like(hg)
    gemm(M, 3 * HD + H, C, [a], w, _lib.EPI_QKV, out=qk2, ldo=2 * HD, q_end=HD, k_end=2 * HD, v_end=3 * HD, q_scale=0.125,
         rope=rope, pos_off=7, rows_per_batch=Nseq, vt=vrows, vt_ld=HD, heads_v=H, hgate=hg2, hgate_ld=H, hgate_bias=hb, v_rowmajor=1)
    assert torch.equal(qk2, qk) and torch.equal(hg2, hg)
    assert rel(vrows[:M], v) < 5e-3 and (vrows[M] == -3.0).all()
    assert torch.equal(vrows[:M].reshape(B, Nseq, H, 64).permute(0, 2, 3, 1).reshape(B * H * 64, Nseq), vt[:, :Nseq])


@pytest.mark.parametrize('H', [4, 8, 16])
def test_qkv_ragged_last_tile_uses_narrow_mma(H):
    """N = 3 H 64 + H: for H = 4 (N = 772) and 16 (N = 3088) the last 256-column tile holds only the H gate columns; its MMA runs
    16 (rounded-up) columns wide.  Every column of the result must still be right."""
    Nseq, B, C = 150, 2, 256
    HD, M = H * 64, B * Nseq
    g = torch.Generator().manual_seed(H)
    a = bf(torch.randn(M, C, generator=g)).to(DEV)
    w = bf(torch.randn(3 * HD + H, C, generator=g) / math.sqrt(C)).to(DEV)
    hb = torch.randn(H, generator=g).to(DEV)
    rope = torch.stack((torch.ones(Nseq, 32), torch.zeros(Nseq, 32)), -1).to(DEV).contiguous()      # identity rotation
    qk = torch.zeros(M, 2 * HD, device=DEV, dtype=torch.bfloat16)
    vrows = torch.zeros(M, HD, device=DEV, dtype=torch.bfloat16)
    hg = torch.zeros(M, H, device=DEV)
    gemm(M, 3 * HD + H, C, [a], w, _lib.EPI_QKV, out=qk, ldo=2 * HD, q_end=HD, k_end=2 * HD, v_end=3 * HD, q_scale=1.0, rope=rope, pos_off=0,
         rows_per_batch=Nseq, vt=vrows, vt_ld=HD, heads_v=H, hgate=hg, hgate_ld=H, hgate_bias=hb, v_rowmajor=1)
    full = a.float() @ w.float().t()
    assert rel(qk, full[:, :2 * HD]) < 5e-3 and rel(vrows, full[:, 2 * HD:3 * HD]) < 5e-3
    assert rel(hg, torch.sigmoid(full[:, 3 * HD:] + hb)) < 1e-4
