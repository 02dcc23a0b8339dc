"""Deterministic synthetic workload: per-clip conditions and self-contained random-init weights with the
reference's parameter names, shapes and init distributions.  Used by bench.py (inputs of the measured path), by the tests and by
the golden-vector scripts (through oracle/synth.py).  Data generators only: nothing here computes any part of the sampling path.

All tensors are fp32 on the host, generated per *global clip index* so that
results are invariant to how clips are sharded over ranks.
"""
from __future__ import annotations

import math

import torch

NOTES = 51            # e2_tts_crossatt3.py:70
FRAME_RATE = 75.0     # 24000 / 320 (torch_tools.py:32-40)
CLIP_FRAMES = 300     # synthetic video frame count per clip
HOP, SR = 320, 24000


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def clip_condition(i: int, n: int, dim_text: int = 1280, F: int = CLIP_FRAMES) -> torch.Tensor:
    """[n, dim_text]: F per-frame embeddings expanded with the reference's nearest-frame rule
    j = min(round((k*320+160)/24000 / (dur/(F-1))), F-1)  (e2_tts_crossatt3.py:1803-1808)."""
    emb = torch.randn(F, dim_text, generator=_gen(1000 + i))
    dur = n / FRAME_RATE
    step = dur / (F - 1)
    idx = [min(int(round(((k * HOP + HOP // 2) / SR) / step)), F - 1) for k in range(n)]
    return emb[torch.tensor(idx)]


def t5_context(i: int, nc: int = 8, dim: int = 1024) -> torch.Tensor:
    return torch.randn(nc, dim, generator=_gen(2000 + i))


def roll_frames(i: int, n: int, live: bool) -> torch.Tensor:
    """V2A: zeros [n, 51] (e2_tts_crossatt3.py:2165).  V2P: sigmoid(N(0,1)*2-2) rows repeated x3, cut/padded to n
    (mirrors e2_tts_crossatt3.py:1545-1554)."""
    if not live:
        return torch.zeros(n, NOTES)
    t = n // 3 + 1
    r = torch.sigmoid(torch.randn(t, NOTES, generator=_gen(3000 + i)) * 2 - 2)
    r = r.repeat_interleave(3, dim=0)
    if r.shape[0] >= n:
        return r[:n]
    return torch.cat([r, torch.zeros(n - r.shape[0], NOTES)], 0)


def noise(i: int, n: int, d: int = 128) -> torch.Tensor:
    return torch.randn(n, d, generator=_gen(4000 + i))


def audio_condition(i: int, n: int, d: int = 128) -> torch.Tensor:
    """Latent frames of a clip's existing audio (the in-painting condition, X3:2224-2228)."""
    return torch.randn(n, d, generator=_gen(6000 + i))


def grey_frames(i: int, F: int, w: int = 100, h: int = 900) -> torch.Tensor:
    """[F, w, h, 1] grey-scale frames as the reference caches them for the piano-roll net (X3:1899-1905)."""
    return torch.rand(F, w, h, 1, generator=_gen(7000 + i))


class StandInRollNet(torch.nn.Module):
    """Deterministic stand-in for the (third-party, pretrained) Video2RollNet in front-end tests: [N, 5, w, h] -> [N, 51]
    logits that depend on every one of the 5 frames of a window and on position inside the frame."""

    def __init__(self, seed: int = 0):
        super().__init__()
        g = _gen(8000 + seed)
        self.register_buffer('mix', torch.randn(5 * 4, NOTES, generator=g))

    def forward(self, x):
        n, _, w, h = x.shape
        rows = torch.tensor([3 % w, 47 % w, 61 % w, w - 1], device=x.device)
        cols = torch.tensor([5 % h, 300 % h, 640 % h, h - 1], device=x.device)
        q = x[:, :, rows, cols]                                                      # [N, 5, 4] probe pixels
        return (q.reshape(n, 20) - 0.5) @ self.mix.to(x.device) * 2.0


def waveform(i: int, nw: int = 240000) -> torch.Tensor:
    return torch.rand(nw, generator=_gen(5000 + i)) - 0.5


def batch(indices, n, *, lens=None, nc=8, dim_text=1280, dim=1024, d=128, live_frames=False, nc_list=None):
    """Stack per-clip conditions into a batch dict.  `lens[i]` (<= n) gives ragged batches; rows past a clip's length
    are zero.  `nc_list` gives ragged T5 lengths (padded to max, mask accordingly)."""
    B = len(indices)
    lens = [n] * B if lens is None else list(lens)
    nc_list = [nc] * B if nc_list is None else list(nc_list)
    ncm = max(nc_list)
    out = dict(
        y0=torch.zeros(B, n, d), clip=torch.zeros(B, n, dim_text), frames=torch.zeros(B, n, NOTES),
        ctx=torch.zeros(B, ncm, dim), ctx_mask=torch.zeros(B, ncm, dtype=torch.bool),
        lens=torch.tensor(lens, dtype=torch.long),
    )
    for b, i in enumerate(indices):
        L = lens[b]
        out['y0'][b] = noise(i, n, d)
        out['clip'][b, :L] = clip_condition(i, L, dim_text)
        out['frames'][b, :L] = roll_frames(i, L, live_frames)
        out['ctx'][b, :nc_list[b]] = t5_context(i, nc_list[b], dim)
        out['ctx_mask'][b, :nc_list[b]] = True
    return out


ZERO_INIT_SUFFIXES = ('to_gamma.weight', 'text_frames_to_audio.weight', 'audio_to_text.weight',
                      'audio_to_frames.weight', 'to_v_head_gate.weight')


def rerandomise_zero_init_(state_dict, seed: int = 1, std: float = 0.02):
    """Default init makes time / CLIP / roll conditioning exactly dead (AdaptiveRMSNorm.to_gamma, AdaLNZero weights,
    the three cross-condition linears and the value-head gate are zero-initialised: e2_tts_crossatt3.py:543-544, 675,
    681, 684 and x-transformers).  Re-draw those groups N(0, std^2) and un-saturate the head gate (bias 10 -> 0) so a
    random-init parity run exercises every path.  In place; returns the dict."""
    g = _gen(seed)
    for k in sorted(state_dict):
        v = state_dict[k]
        if k.endswith(ZERO_INIT_SUFFIXES):
            v.copy_(torch.randn(v.shape, generator=g) * std)
        elif k.endswith('to_v_head_gate.bias'):
            v.zero_()
    return state_dict


# ----------------------------------------------------------------------------
# self-contained random-init weights with the reference's names, shapes and init distributions
# ----------------------------------------------------------------------------

def random_state_dict(*, depth=12, dim=1024, dim_text=1280, dim_frames=512, heads=16, dim_head=64, frames_heads=8,
                      ff_mult=4, num_channels=128, max_seq_len=8192, num_registers=32, kernel_size=31, seed=0,
                      live_conditioning=True, cond_proj_in=False):
    """State dict of the sampling path (SURVEY.md Appendix C key names) drawn from a seeded CPU generator.

    Mirrors the distributions the reference constructor would give (nn.Linear / nn.Conv1d default
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)); registers N(0, 0.02^2) X3:767-775; nn.Embedding N(0,1); RMSNorm g = 1;
    AdaLNZero bias -2 X3:544; RandomFourierEmbed N(0,1) X3:559) and, with `live_conditioning`, applies
    `rerandomise_zero_init_` so no conditioning path is dead.  It exists because the reference constructor cannot
    travel to the GPU box; tests/test_oracle_vs_reference.py checks names and shapes against the real constructor.
    """
    g = _gen(10_000 + seed)
    sd = {}

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    def linear(name, out_f, in_f, bias=True):
        sd[name + '.weight'] = uni((out_f, in_f), in_f)
        if bias:
            sd[name + '.bias'] = uni((out_f,), in_f)

    def attn(name, d_model, h):
        inner = h * dim_head
        linear(name + '.to_q', inner, d_model, bias=False)
        linear(name + '.to_k', inner, d_model, bias=False)
        linear(name + '.to_v', inner, d_model, bias=False)
        sd[name + '.to_v_head_gate.weight'] = torch.zeros(h, d_model)
        sd[name + '.to_v_head_gate.bias'] = torch.full((h,), 10.0)
        linear(name + '.to_out', d_model, inner, bias=False)

    def ff(name, d_model):
        inner = d_model * ff_mult
        linear(name + '.ff.0.proj', inner * 2, d_model)
        linear(name + '.ff.2', d_model, inner)

    def conv(name, c):
        sd[name + '.dw_conv1d.0.weight'] = uni((c, 1, kernel_size), kernel_size)
        sd[name + '.dw_conv1d.0.bias'] = uni((c,), kernel_size)

    t = 'transformer.'
    sd[t + 'registers'] = torch.randn(num_registers, dim, generator=g) * 0.02
    sd[t + 'text_registers'] = torch.randn(num_registers, dim_text, generator=g) * 0.02
    sd[t + 'frames_registers'] = torch.randn(num_registers, dim_frames, generator=g) * 0.02
    sd[t + 'abs_pos_emb.weight'] = torch.randn(max_seq_len, dim, generator=g)
    inv_freq = 1. / (10000 ** (torch.arange(0, dim_head, 2).float() / dim_head))
    for r in ('rotary_emb', 'text_rotary_emb', 'frames_rotary_emb'):
        sd[t + r + '.inv_freq'] = inv_freq.clone()
    sd[t + 'time_cond_mlp.0.weights'] = torch.randn(dim // 2, generator=g)
    linear(t + 'time_cond_mlp.1', dim, dim + 1)
    sd[t + 'final_norm.g'] = torch.ones(dim)

    for L in range(depth):
        a = f'{t}layers.{L}.0.'
        if L >= depth // 2:
            linear(a + '0', dim, dim * 2, bias=False)
        conv(a + '1', dim)
        for norm_i, attn_i, ada_i in ((2, 3, 4), (5, 6, 7)):
            sd[a + f'{norm_i}.to_gamma.weight'] = torch.zeros(dim, dim)
            attn(a + f'{attn_i}', dim, heads)
            sd[a + f'{ada_i}.to_gamma.weight'] = torch.zeros(dim, dim)
            sd[a + f'{ada_i}.to_gamma.bias'] = torch.full((dim,), -2.0)
        sd[a + '8.to_gamma.weight'] = torch.zeros(dim, dim)
        ff(a + '9', dim)
        sd[a + '10.to_gamma.weight'] = torch.zeros(dim, dim)
        sd[a + '10.to_gamma.bias'] = torch.full((dim,), -2.0)

        x = f'{t}layers.{L}.1.'
        conv(x + '0', dim_text)
        sd[x + '1.g'] = torch.ones(dim_text)
        attn(x + '2', dim_text, heads)
        sd[x + '3.g'] = torch.ones(dim_text)
        ff(x + '4', dim_text)
        sd[x + '5.text_frames_to_audio.weight'] = torch.zeros(dim, dim + dim_text + dim_frames)
        if L < depth - 1:
            sd[x + '5.audio_to_text.weight'] = torch.zeros(dim_text, dim + dim_text)
            sd[x + '5.audio_to_frames.weight'] = torch.zeros(dim_frames, dim + dim_frames)

        f = f'{t}layers.{L}.2.'
        conv(f + '0', dim_frames)
        sd[f + '1.g'] = torch.ones(dim_frames)
        attn(f + '2', dim_frames, frames_heads)
        sd[f + '3.g'] = torch.ones(dim_frames)
        ff(f + '4', dim_frames)

    linear('proj_in', dim, num_channels)
    linear('to_pred', num_channels, dim)
    linear('proj_frames', dim_frames, NOTES)
    if cond_proj_in:                          # E2TTS(if_cond_proj_in=True): the audio-condition projection, X3:1365 (drawn last:
        linear('cond_proj_in', dim, num_channels)   # the other tensors do not depend on this switch)
    if live_conditioning:
        rerandomise_zero_init_(sd, seed=seed + 1)
    return sd


TINY = dict(depth=4, dim=128, dim_text=192, dim_frames=64, heads=2, dim_head=64, num_channels=64, max_seq_len=256)
SHIPPED = dict(depth=12, dim=1024, dim_text=1280, dim_frames=512, heads=16, dim_head=64, num_channels=128,
               max_seq_len=8192)
