"""TEST INFRASTRUCTURE ONLY -- restated third-party math for the oracle.

The arithmetic of the reference hot path lives in pinned dependencies that are
not vendored under /root/reference and are not installable here (no network):

  * x-transformers==1.37.4  (requirements.txt:19; imported at
    src/e2_tts_pytorch/e2_tts_crossatt3.py:38-45) -- Attention, FeedForward,
    RMSNorm, AdaptiveRMSNorm, RotaryEmbedding
  * torchdiffeq==0.2.4      (requirements.txt:12; e2_tts_crossatt3.py:32,2255)
  * einx==0.3.0             (requirements.txt:4;  e2_tts_crossatt3.py:305,314,
    347-351,519-526,562,658)

This module restates the published algorithm of exactly the symbols the
reference calls, with the same constructor signatures / parameter names so the
reference file imports and runs against it (see oracle/ref_loader.py), and so
state-dict keys match SURVEY.md Appendix C.

PARITY STATUS: *unpinned by the reference* -- the reference ships no tests and
no golden vectors.  The restatement cannot be diffed against
a real install offline; tests/test_oracle_third_party.py pins it against
hand-derived values instead.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this package.  The product path never does.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


# ----------------------------------------------------------------------------
# x-transformers 1.37.4
# ----------------------------------------------------------------------------

class RMSNorm(nn.Module):
    """y = normalize(x) * sqrt(dim) * g   (F.normalize eps = 1e-12)."""

    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.g = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return F.normalize(x, dim=-1) * self.scale * self.g


class AdaptiveRMSNorm(nn.Module):
    """y = normalize(x) * sqrt(dim) * (to_gamma(cond) + 1); to_gamma zero-init, no bias."""

    def __init__(self, dim, dim_condition=None):
        super().__init__()
        self.scale = dim ** 0.5
        dim_condition = dim if dim_condition is None else dim_condition
        self.to_gamma = nn.Linear(dim_condition, dim, bias=False)
        nn.init.zeros_(self.to_gamma.weight)

    def forward(self, x, *, condition):
        if condition.ndim == 2:
            condition = condition[:, None, :]
        normed = F.normalize(x, dim=-1)
        gamma = self.to_gamma(condition)
        return normed * self.scale * (gamma + 1.)


class _GLU(nn.Module):
    """proj -> chunk(2): first half is the value, second half the gate; value * gelu_erf(gate)."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    """FeedForward(dim, glu=True, mult=4, dropout): ff = Sequential(GLU, Dropout, Linear)."""

    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0., **kwargs):
        super().__init__()
        assert not kwargs, f'oracle FeedForward: unsupported kwargs {kwargs}'
        inner = int(dim * mult)
        dim_out = dim if dim_out is None else dim_out
        if glu:
            project_in = _GLU(dim, inner)
        else:
            project_in = nn.Sequential(nn.Linear(dim, inner), nn.GELU())
        self.ff = nn.Sequential(project_in, nn.Dropout(dropout), nn.Linear(inner, dim_out))

    def forward(self, x):
        return self.ff(x)


class RotaryEmbedding(nn.Module):
    """inv_freq[j] = 10000^(-2j/dim); freqs[pos, 2j] = freqs[pos, 2j+1] = pos * inv_freq[j]."""

    def __init__(self, dim, base=10000):
        super().__init__()
        inv_freq = 1. / (base ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer('inv_freq', inv_freq)

    def forward_from_seq_len(self, seq_len):
        t = torch.arange(seq_len, device=self.inv_freq.device)
        return self.forward(t)

    def forward(self, t):
        if t.ndim == 1:
            t = t[None, :]
        freqs = torch.einsum('b i , j -> b i j', t.type_as(self.inv_freq), self.inv_freq)
        freqs = torch.stack((freqs, freqs), dim=-1).flatten(-2)   # interleaved pairs
        return freqs, 1.


def rotate_half(x):
    """adjacent pairs (x[2j], x[2j+1]) -> (-x[2j+1], x[2j])."""
    x = x.unflatten(-1, (-1, 2))
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def apply_rotary_pos_emb(t, freqs, scale=1.):
    """Uses the LAST seq_len rows of the table (so cross-attention keys of nc rows get positions N-nc..N-1)."""
    rot_dim, seq_len, orig_dtype = freqs.shape[-1], t.shape[-2], t.dtype
    freqs = freqs[:, -seq_len:, :]
    if t.ndim == 4 and freqs.ndim == 3:
        freqs = freqs[:, None, :, :]
    t, t_unrotated = t[..., :rot_dim], t[..., rot_dim:]
    t = (t * freqs.cos() * scale) + (rotate_half(t) * freqs.sin() * scale)
    return torch.cat((t, t_unrotated), dim=-1).type(orig_dtype)


class Attention(nn.Module):
    """Attention(dim, heads, dim_head, dropout, gate_value_heads=True, softclamp_logits=True).

    sim = (q k^T) * dim_head^-0.5 ; sim = tanh(sim/50)*50 ; key-padding mask ->
    -finfo.max ; softmax in fp32 ; out = attn v ; per-(token, head) sigmoid gate
    from the query-side input ; to_out ; zero padded query rows.
    """

    def __init__(self, dim, heads=8, dim_head=64, dropout=0., gate_value_heads=False,
                 softclamp_logits=False, logit_softclamp_value=50., dim_context=None, **kwargs):
        super().__init__()
        assert not kwargs, f'oracle Attention: unsupported kwargs {kwargs}'
        dim_kv = dim if dim_context is None else dim_context
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head ** -0.5
        self.softclamp_logits = softclamp_logits
        self.logit_softclamp_value = logit_softclamp_value
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim_kv, inner, bias=False)
        self.to_v = nn.Linear(dim_kv, inner, bias=False)
        self.to_v_head_gate = None
        if gate_value_heads:
            self.to_v_head_gate = nn.Linear(dim, heads)
            nn.init.constant_(self.to_v_head_gate.weight, 0)
            nn.init.constant_(self.to_v_head_gate.bias, 10)
        self.dropout = nn.Dropout(dropout)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, context=None, mask=None, context_mask=None, rotary_pos_emb=None):
        b, n, h, d = x.shape[0], x.shape[1], self.heads, self.dim_head
        has_context = context is not None
        kv_input = context if has_context else x

        def split(t):
            return t.unflatten(-1, (h, d)).transpose(1, 2)   # b n (h d) -> b h n d

        q, k, v = split(self.to_q(x)), split(self.to_k(kv_input)), split(self.to_v(kv_input))

        if rotary_pos_emb is not None:
            freqs, _ = rotary_pos_emb
            q = apply_rotary_pos_emb(q, freqs)
            k = apply_rotary_pos_emb(k, freqs)

        input_mask = context_mask
        if input_mask is None and not has_context:
            input_mask = mask

        sim = torch.einsum('b h i d, b h j d -> b h i j', q, k) * self.scale
        if self.softclamp_logits:
            sim = (sim / self.logit_softclamp_value).tanh() * self.logit_softclamp_value
        if input_mask is not None:
            sim = sim.masked_fill(~input_mask[:, None, None, :], -torch.finfo(sim.dtype).max)
        attn = F.softmax(sim, dim=-1, dtype=torch.float32).type(sim.dtype)
        attn = self.dropout(attn)
        out = torch.einsum('b h i j, b h j d -> b h i d', attn, v)

        if self.to_v_head_gate is not None:
            head_gate = self.to_v_head_gate(x)                       # b n h
            out = out * head_gate.sigmoid().transpose(1, 2)[..., None]

        out = out.transpose(1, 2).flatten(-2)                        # b n (h d)
        out = self.to_out(out)
        if mask is not None:
            out = out.masked_fill(~mask[:, :, None], 0.)
        return out


# ----------------------------------------------------------------------------
# torchdiffeq 0.2.4 -- fixed-grid Euler on the caller's grid
# ----------------------------------------------------------------------------

def odeint(fn, y0, t, *, method='euler', **kwargs):
    """y[i+1] = y[i] + (t[i+1]-t[i]) * fn(t[i], y[i]); returns stack(y[0..len(t)-1])."""
    assert method == 'euler', 'oracle odeint restates only the fixed-grid Euler solver'
    ys = [y0]
    y = y0
    for i in range(len(t) - 1):
        t0, t1 = t[i], t[i + 1]
        dt = t1 - t0
        y = y + dt * fn(t0.to(y.dtype), y)
        ys.append(y)
    return torch.stack(ys)


# ----------------------------------------------------------------------------
# einx 0.3.0 -- exactly the patterns the reference uses
# ----------------------------------------------------------------------------

class _Einx:
    @staticmethod
    def less(pattern, a, b):
        assert pattern == 'n, b -> b n', pattern
        return a[None, :] < b[:, None]

    @staticmethod
    def greater_equal(pattern, a, b):
        assert pattern == 'n, b -> b n', pattern
        return a[None, :] >= b[:, None]

    @staticmethod
    def where(pattern, m, x, y):
        if pattern == 'b n, b n d, -> b n d':
            return torch.where(m[..., None], x, torch.as_tensor(y, dtype=x.dtype, device=x.device))
        if pattern == 'b n, b n d, b n d -> b n d':
            return torch.where(m[..., None], x, y)
        raise AssertionError(pattern)

    @staticmethod
    def multiply(pattern, a, b):
        if pattern == 'i, j -> i j':
            return a[:, None] * b[None, :]
        if pattern == 'b n h, b h n d -> b h n d':
            return a.transpose(1, 2)[..., None] * b
        raise AssertionError(pattern)

    @staticmethod
    def divide(pattern, a, b):
        assert pattern == 'b d, b -> b d', pattern
        return a / b[:, None]


einx = _Einx()
