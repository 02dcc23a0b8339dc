"""Pins for the restated third-party arithmetic (oracle/third_party.py) against hand-derived values.

The reference has no tests; x-transformers 1.37.4 / torchdiffeq 0.2.4 / einx 0.3.0 are not
installable offline, so these known-answer cases are the pin for the restatement.
"""
import math

import torch

from oracle import third_party as tp


def test_rotary_is_interleaved_pairs():
    rot = tp.RotaryEmbedding(64)
    freqs, scale = rot.forward_from_seq_len(5)
    assert scale == 1.
    assert freqs.shape == (1, 5, 64)
    inv = 1. / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    assert torch.equal(freqs[0, 3, 0::2], 3 * inv) and torch.equal(freqs[0, 3, 1::2], 3 * inv)


def test_rotate_half_adjacent_pairs():
    x = torch.tensor([1., 2., 3., 4.])
    assert torch.equal(tp.rotate_half(x), torch.tensor([-2., 1., -4., 3.]))


def test_rope_uses_last_rows_and_rotates_pairs():
    freqs = torch.zeros(1, 6, 4)
    freqs[0, :, :] = torch.arange(6.)[:, None] * math.pi / 2      # position p rotates every pair by p*90deg
    t = torch.tensor([1., 0., 0., 1.]).expand(1, 1, 2, 4)          # 2 rows -> positions 4 and 5
    out = tp.apply_rotary_pos_emb(t, freqs)
    # position 4: 360deg -> identity; position 5: 450deg = 90deg: (1,0)->(0,1), (0,1)->(-1,0)
    assert torch.allclose(out[0, 0, 0], torch.tensor([1., 0., 0., 1.]), atol=1e-5)
    assert torch.allclose(out[0, 0, 1], torch.tensor([0., 1., -1., 0.]), atol=1e-5)


def test_rmsnorm_and_adaptive():
    x = torch.tensor([[[3., 4.]]])
    n = tp.RMSNorm(2)
    assert torch.allclose(n(x), torch.tensor([[[0.6, 0.8]]]) * math.sqrt(2))
    a = tp.AdaptiveRMSNorm(2)
    assert torch.count_nonzero(a.to_gamma.weight) == 0
    with torch.no_grad():
        a.to_gamma.weight.copy_(torch.tensor([[1., 0.], [0., 2.]]))
    cond = torch.tensor([[0.5, 0.25]])
    assert torch.allclose(a(x, condition=cond), torch.tensor([[[0.6 * 1.5, 0.8 * 1.5]]]) * math.sqrt(2))
    assert torch.equal(n(torch.zeros(1, 1, 2)), torch.zeros(1, 1, 2))      # eps path: 0 / max(0, 1e-12)


def test_geglu_value_first_gate_second_and_keys():
    ff = tp.FeedForward(dim=2, glu=True, mult=1, dropout=0.)
    assert set(ff.state_dict()) == {'ff.0.proj.weight', 'ff.0.proj.bias', 'ff.2.weight', 'ff.2.bias'}
    with torch.no_grad():
        ff.ff[0].proj.weight.copy_(torch.tensor([[1., 0.], [0., 1.], [2., 0.], [0., 2.]]))
        ff.ff[0].proj.bias.zero_()
        ff.ff[2].weight.copy_(torch.eye(2)); ff.ff[2].bias.zero_()
    x = torch.tensor([[0.5, -1.0]])
    gelu = lambda v: 0.5 * v * (1 + math.erf(v / math.sqrt(2)))
    want = torch.tensor([[0.5 * gelu(1.0), -1.0 * gelu(-2.0)]])
    assert torch.allclose(ff(x), want, atol=1e-6)


def test_attention_softclamp_gate_mask():
    torch.manual_seed(0)
    att = tp.Attention(dim=64, heads=1, dim_head=64, gate_value_heads=True, softclamp_logits=True)
    assert set(att.state_dict()) == {'to_q.weight', 'to_k.weight', 'to_v.weight', 'to_v_head_gate.weight',
                                     'to_v_head_gate.bias', 'to_out.weight'}
    assert torch.all(att.to_v_head_gate.bias == 10) and torch.count_nonzero(att.to_v_head_gate.weight) == 0
    with torch.no_grad():
        for lin in (att.to_q, att.to_k, att.to_v, att.to_out):
            lin.weight.copy_(torch.eye(64))
        att.to_v_head_gate.bias.zero_()
    x = torch.zeros(1, 3, 64)
    x[0, 0, 0] = 100.; x[0, 1, 0] = 100.; x[0, 2, 1] = 7.
    mask = torch.tensor([[True, True, False]])
    out = att(x, mask=mask)
    # logits row0: q0.k0 = 1e4/8 = 1250 -> clamp 50*tanh(25) = 50 (same for k1); key 2 masked -> equal weights .5/.5
    # gate = sigmoid(0) = .5 ; out row0 = .5 * (.5*v0 + .5*v1) = .5 * 100 e0 ; row 2 (padded query) zeroed
    assert torch.allclose(out[0, 0, 0], torch.tensor(50.), atol=1e-4)
    assert torch.equal(out[0, 2], torch.zeros(64))
    # unsaturated case: q.k/8 = 2 vs 0 -> softclamped logits 50*tanh(2/50), 0
    x2 = torch.zeros(1, 2, 64); x2[0, 0, 0] = 4.; x2[0, 1, 1] = 4.
    o2 = att(x2)
    s = 50 * math.tanh(2 / 50)
    p = math.exp(s) / (math.exp(s) + 1)
    assert torch.allclose(o2[0, 0, 0], torch.tensor(0.5 * p * 4.), atol=1e-5)
    assert torch.allclose(o2[0, 0, 1], torch.tensor(0.5 * (1 - p) * 4.), atol=1e-5)


def test_cross_attention_null_context_is_exact_zero():
    torch.manual_seed(0)
    att = tp.Attention(dim=64, heads=1, dim_head=64, gate_value_heads=True, softclamp_logits=True)
    rot = tp.RotaryEmbedding(64).forward_from_seq_len(10)
    x = torch.randn(2, 10, 64)
    out = att(x, context=torch.zeros(2, 4, 64), context_mask=torch.ones(2, 4, dtype=torch.bool), rotary_pos_emb=rot)
    assert torch.count_nonzero(out) == 0


def test_euler_on_given_grid():
    t = torch.tensor([0., 0.1, 0.4, 1.0])
    seen = []
    def fn(tt, y):
        seen.append(float(tt)); return torch.full_like(y, 2.0) * tt
    ys = tp.odeint(fn, torch.zeros(2), t, method='euler')
    assert ys.shape == (4, 2) and [round(s, 6) for s in seen] == [0., 0.1, 0.4]
    # y1 = 0 ; y2 = 0 + .3*(2*.1) = .06 ; y3 = .06 + .6*(2*.4) = .54
    assert torch.allclose(ys[:, 0], torch.tensor([0., 0., 0.06, 0.54]), atol=1e-6)


def test_einx_patterns():
    e = tp.einx
    assert torch.equal(e.less('n, b -> b n', torch.arange(3), torch.tensor([1, 3])),
                       torch.tensor([[True, False, False], [True, True, True]]))
    m = torch.tensor([[True, False]])
    x = torch.ones(1, 2, 2)
    assert torch.equal(e.where('b n, b n d, -> b n d', m, x, 0.), torch.tensor([[[1., 1.], [0., 0.]]]))
    assert torch.equal(e.multiply('i, j -> i j', torch.tensor([1., 2.]), torch.tensor([3., 4.])),
                       torch.tensor([[3., 4.], [6., 8.]]))
