"""Error-compensated "fp32" mode (precision='fp32'): rel-L2 <= 1e-4 against the fp32 reference (north_star tolerance)."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import DEV, L, gemm, kcheck, rel
from e2_tts_pytorch import _lib
from oracle import synth
from test_gpu_4_path import build_model, dev, load_gold, valid_rel

TOL32 = 1e-4


def _split(t):
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def test_split_gemm_is_fp32_accurate():
    """A (hi|lo) x W (hi|hi|lo) through the unchanged tcgen05 kernel: 3 MMAs per product, fp32 accumulate."""
    M, N, K = 500, 256, 512
    g = torch.Generator().manual_seed(0)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    ah, al = _split(a)
    wh, wl = _split(w)
    a2 = torch.cat([ah, al], 1).contiguous()                    # [M, 2K]
    w3 = torch.cat([wh, wh, wl], 1).contiguous()                # [N, 3K]
    out = torch.zeros(M, N, device=DEV)
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.num_src = M, N, 3 * K, 3
    for i, off in enumerate((0, K, 0)):
        d.a[i] = a2.data_ptr() + off * 2; d.lda[i] = 2 * K; d.ka[i] = K
    d.w = w3.data_ptr(); d.ldw = 3 * K; d.epi = _lib.EPI_F32; d.out = out.data_ptr(); d.ldo = N
    kcheck(L().e2b_gemm_launch(C.byref(d), _lib.stream_ptr()))
    ref = (a.double() @ w.double().t()).float()
    e = rel(out, ref)
    print(f'split GEMM rel {e:.3e} (plain bf16 would be ~2e-3)')
    assert e < 2e-5


def test_split_store_epilogue():
    M, N, K = 300, 128, 64
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / 8).to(torch.bfloat16)
    out = torch.zeros(M, 2 * N, device=DEV, dtype=torch.bfloat16)
    gemm(M, N, K, [a], w, _lib.EPI_BF16, out=out, ldo=2 * N, split=N)
    ref = a.float() @ w.float().t()
    assert rel(out[:, :N].float() + out[:, N:].float(), ref) < 2e-5
    assert rel(out[:, :N], ref) < 4e-3 and rel(out[:, :N], ref) > 1e-4       # the hi half alone is only bf16-accurate


def test_attention_f32_kernel():
    B, H, N = 2, 3, 200
    g = torch.Generator().manual_seed(1)
    q = (torch.randn(B * N, H * 64, generator=g) * 0.4).to(DEV)
    k = (torch.randn(B * N, H * 64, generator=g) * 2).to(DEV)
    v = torch.randn(B * N, H * 64, generator=g).to(DEV)
    gate = torch.rand(B * N, H, generator=g).to(DEV)
    lens = torch.tensor([200, 131], device=DEV, dtype=torch.int32)
    out = torch.zeros(B * N, 2 * H * 64, device=DEV, dtype=torch.bfloat16)
    d = _lib.AttnF32Desc(batch=B, heads=H, q_rows_per_batch=N, kv_rows_per_batch=N, q=q.data_ptr(), ldq=H * 64, q_col0=0, k=k.data_ptr(),
                         ldk=H * 64, k_col0=0, v=v.data_ptr(), ldv=H * 64, v_col0=0, kv_batch_mod=0, kv_lens=lens.data_ptr(), kv_lens_add=0,
                         hgate=gate.data_ptr(), hgate_ld=H, out=out.data_ptr(), ldo=2 * H * 64, out_split=H * 64, softclamp=50.0)
    kcheck(L().e2b_attention_f32_launch(C.byref(d), _lib.stream_ptr()))
    sp = lambda t: t.reshape(B, N, H, 64).permute(0, 2, 1, 3).double()
    sim = torch.einsum('bhid,bhjd->bhij', sp(q), sp(k))
    sim = torch.tanh(sim / 50) * 50
    mask = torch.arange(N, device=DEV)[None, :] < lens[:, None]
    sim = sim.masked_fill(~mask[:, None, None, :], -1e300)
    ref = torch.einsum('bhij,bhjd->bhid', sim.softmax(-1), sp(v)) * gate.reshape(B, N, H).permute(0, 2, 1)[..., None]
    ref = ref.permute(0, 2, 1, 3).reshape(B * N, H * 64).float()
    got = out[:, :H * 64].float() + out[:, H * 64:].float()
    assert rel(got, ref) < 2e-5


def test_tiny_fp32_mode_vs_x3_golden():
    g, r, cfg, bt = load_gold('tiny_x3.pt')
    m, _ = build_model(cfg, r['weight_seed'])
    m.precision = 'fp32'
    d = dev(bt)
    pred = m.velocity(d['y0'], r['t_single'], clip=d['clip'], context=d['ctx'], context_mask=d['ctx_mask'], roll=d['frames'],
                      lens=r['lens'], passes=('null',))
    e0, e1 = valid_rel(pred[0], g['pred_cond'], r['lens']), valid_rel(pred[1], g['pred_null'], r['lens'])
    out = m.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=d['lens'], duration=d['lens'], steps=r['steps'],
                   cfg_strength=r['cfg_strength'], remove_parallel_component=False, return_raw_output=True, context=d['ctx'],
                   context_mask=d['ctx_mask'], frames=d['frames'], noise=d['y0'])
    e2 = valid_rel(out, g['sample_cfg'], r['lens'])
    print(f'tiny fp32 mode rel-L2: cond {e0:.3e} null {e1:.3e} sample {e2:.3e}')
    assert max(e0, e1, e2) < TOL32


@pytest.mark.timeout(900)
def test_shipped_fp32_mode_vs_x3_golden():
    g, r, cfg, bt = load_gold('shipped_x3.pt')
    m, _ = build_model(cfg, r['weight_seed'])
    m.precision = 'fp32'
    d = dev(bt)
    pred = m.velocity(d['y0'], r['t_single'], clip=d['clip'], context=d['ctx'], context_mask=d['ctx_mask'], roll=None, lens=r['lens'],
                      passes=('null',))
    e0, e1 = rel(pred[0], g['pred_cond'].to(DEV)), rel(pred[1], g['pred_null'].to(DEV))
    out = m.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=d['lens'], duration=d['lens'], steps=r['steps'],
                   cfg_strength=r['cfg_strength'], remove_parallel_component=False, return_raw_output=True, context=d['ctx'],
                   context_mask=d['ctx_mask'], noise=d['y0'])
    e2 = rel(out, g['sample_cfg'].to(DEV))
    print(f'shipped arch fp32 mode rel-L2: cond {e0:.3e} null {e1:.3e} sample {e2:.3e}')
    assert max(e0, e1, e2) < TOL32
