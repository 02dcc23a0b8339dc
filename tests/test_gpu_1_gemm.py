"""tcgen05 GEMM (csrc/gemm.cu) against torch on the same bf16 operands."""
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
