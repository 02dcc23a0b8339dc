"""tcgen05 flash-style attention (csrc/attention.cu) against a materialised fp32 torch reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import DEV, attention, bf, rel


def _ref(q, k, v, kv_lens, gate, clamp=50.0):
    # q [B,H,Nq,64] (pre-scaled), k [B,H,Nk,64], v [B,H,Nk,64]
    sim = torch.einsum('bhid,bhjd->bhij', q.float(), k.float())
    sim = torch.tanh(sim / clamp) * clamp
    Nk = k.shape[2]
    mask = torch.arange(Nk, device=q.device)[None, :] < kv_lens[:, None]
    sim = sim.masked_fill(~mask[:, None, None, :], -torch.finfo(torch.float32).max)
    out = torch.einsum('bhij,bhjd->bhid', sim.softmax(-1), v.float())
    return out * gate.permute(0, 2, 1)[..., None]


# the last case has more kv sequences than the kernel stages in shared memory (ATT_MAXB = 1024): lengths come from global memory
def _v_operand(v, N, rows):
    """V [B,H,N,64] in the layout the kernel is asked for: plain rows [B*N, H*64] (MN-major operand) or V^T [(b,h,d), keys]."""
    B, H = v.shape[:2]
    if rows:
        return v.permute(0, 2, 1, 3).reshape(B * N, H * 64).contiguous(), H * 64
    Npad = (N + 7) // 8 * 8
    vt = torch.zeros(B * H * 64, Npad, device=DEV, dtype=torch.bfloat16)
    vt[:, :N] = v.permute(0, 1, 3, 2).reshape(B * H * 64, N)
    return vt, Npad


# the last case has more kv sequences than the kernel stages in shared memory (ATT_MAXB = 1024): lengths come from global memory
@pytest.mark.parametrize('rows', [True, False], ids=['v_rows', 'v_transposed'])
@pytest.mark.parametrize('B,H,N,lens', [(1, 1, 128, [128]), (2, 2, 100, [100, 61]), (2, 3, 300, [300, 257]),
                                        (1, 16, 782, [782]), (3, 8, 782, [782, 500, 129]), (2, 2, 407, [407, 233]), (1, 2, 2282, [2282]),
                                        (1100, 1, 40, [40 - (i % 29) for i in range(1100)])])
def test_self_attention(B, H, N, lens, rows):
    g = torch.Generator().manual_seed(N + H)
    HD = H * 64
    qk = bf(torch.randn(B * N, 2 * HD, generator=g)).to(DEV)
    qk[:, :HD] *= 0.35       # plays the role of the 1/8 pre-scale with larger logits to exercise the soft-clamp
    qk[:, HD:] *= 3.0
    v = bf(torch.randn(B, H, N, 64, generator=g)).to(DEV)
    vt, Npad = _v_operand(v, N, rows)
    gate = torch.rand(B, N, H, generator=g).to(DEV)
    out = torch.full((B * N, HD), float('nan'), device=DEV, dtype=torch.bfloat16)
    kv = torch.tensor(lens, device=DEV, dtype=torch.int32)
    attention(batch=B, heads=H, q_rows_per_batch=N, kv_rows_per_batch=N, q=qk, ldq=2 * HD, q_col0=0, k=qk, ldk=2 * HD, k_col0=HD,
              vt=vt, vt_ld=Npad, kv_batch_mod=0, kv_lens=kv, kv_lens_add=0, hgate=gate.reshape(B * N, H).contiguous(), hgate_ld=H,
              out=out, ldo=HD, softclamp=50.0, v_rowmajor=int(rows), v_col0=0)
    q = qk[:, :HD].reshape(B, N, H, 64).permute(0, 2, 1, 3)
    k = qk[:, HD:].reshape(B, N, H, 64).permute(0, 2, 1, 3)
    ref = _ref(q, k, v, kv, gate).permute(0, 2, 1, 3).reshape(B * N, HD)
    e = rel(out, ref)
    print(f'attention B{B} H{H} N{N}: rel {e:.3e}, max logit {torch.einsum("bhid,bhjd->bhij", q.float(), k.float()).abs().max().item():.1f}')
    assert torch.isfinite(out.float()).all()
    assert e < 1.5e-2


@pytest.mark.parametrize('rows', [True, False], ids=['v_rows', 'v_transposed'])
def test_cross_attention_shared_context(rows):
    B, P, H, N, nc = 2, 2, 2, 150, 8
    HD = H * 64
    g = torch.Generator().manual_seed(9)
    q = bf(torch.randn(P * B * N, HD, generator=g) * 0.3).to(DEV)
    k = bf(torch.randn(B * nc, HD, generator=g)).to(DEV)
    v = bf(torch.randn(B, H, nc, 64, generator=g)).to(DEV)
    vt, vld = _v_operand(v, nc, rows)
    lens = torch.tensor([8, 5], device=DEV, dtype=torch.int32)
    out = torch.zeros(P * B * N, HD, device=DEV, dtype=torch.bfloat16)
    attention(batch=P * B, heads=H, q_rows_per_batch=N, kv_rows_per_batch=nc, q=q, ldq=HD, q_col0=0, k=k, ldk=HD, k_col0=0, vt=vt,
              vt_ld=vld, kv_batch_mod=B, kv_lens=lens, kv_lens_add=0, hgate=0, hgate_ld=0, out=out, ldo=HD, softclamp=50.0,
              v_rowmajor=int(rows), v_col0=0)
    qq = q.reshape(P * B, N, H, 64).permute(0, 2, 1, 3)
    kk = k.reshape(B, nc, H, 64).permute(0, 2, 1, 3).repeat(P, 1, 1, 1)
    ref = _ref(qq, kk, v.repeat(P, 1, 1, 1), lens.repeat(P), torch.ones(P * B, N, H, device=DEV))
    assert rel(out, ref.permute(0, 2, 1, 3).reshape(P * B * N, HD)) < 1.5e-2
