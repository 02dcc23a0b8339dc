"""ORACLE (test infrastructure, not the product): CPU restatement of the EnCodec decoder the reference calls through
`EncodecWrapper.decode` (src/e2_tts_pytorch/e2_tts_crossatt3.py:434-437: `self.model.decoder(emb)`, `output[0]`).

The algorithm lives in a third-party dependency that is not vendored under /root/reference: HuggingFace **transformers==4.46.0**
(requirements.txt:20), `models/encodec/modeling_encodec.py` -- `EncodecDecoder`, `EncodecConv1d`, `EncodecConvTranspose1d`,
`EncodecLSTM`, `EncodecResnetBlock` with the `facebook/encodec_24khz` configuration (causal, weight-norm, reflect padding,
ratios (8,5,4,2), 32 filters, 2 LSTM layers, ELU).  Restated here over a plain state dict with torch functional ops:

  * weight norm      w = g * v / ||v||, norm over all dims but 0 (torch.nn.utils.parametrizations.weight_norm, dim=0)
  * causal Conv1d    left padding (K-1) in reflect mode, zero-extended first when the signal is not longer than the padding
  * ConvTranspose1d  kernel 2s, stride s, then the causal trim of the right (K - s) samples (trim_right_ratio = 1)
  * LSTM             h, c from zeros; gates i, f, g, o; output = lstm(x) + x
  * ResnetBlock      shortcut_conv1x1(x) + conv1x1(ELU(conv3(ELU(x))))

Pinned against the HuggingFace module itself (transformers is part of the image here and on the GPU box:
tests/test_oracle_encodec.py) and against a committed fixture made with it (tests/golden/encodec_tiny.pt,
oracle/make_golden_encodec.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

DEFAULT = dict(hidden_size=128, num_filters=32, upsampling_ratios=(8, 5, 4, 2), kernel_size=7, last_kernel_size=7,
               residual_kernel_size=3, num_lstm_layers=2, audio_channels=1, compress=2)


def weight(sd, prefix):
    """Effective weight of a weight-normed conv: g * v / ||v|| with the norm over every dim but 0."""
    g, v = sd[prefix + '.conv.parametrizations.weight.original0'], sd[prefix + '.conv.parametrizations.weight.original1']
    return g * v / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.ndim - 1)))


def pad_left_reflect(x, p):
    """HF EncodecConv1d._pad1d for paddings (p, 0): reflect, zero-extending short signals first."""
    if p == 0:
        return x
    length = x.shape[-1]
    extra = 0
    if length <= p:
        extra = p - length + 1
        x = F.pad(x, (0, extra))
    y = F.pad(x, (p, 0), mode='reflect')
    return y[..., :y.shape[-1] - extra]


def conv1d(sd, prefix, x):
    w = weight(sd, prefix)
    return F.conv1d(pad_left_reflect(x, w.shape[-1] - 1), w, sd[prefix + '.conv.bias'])


def conv_transpose1d(sd, prefix, x, stride):
    w = weight(sd, prefix)                                   # [Ci, Co, K]
    y = F.conv_transpose1d(x, w, sd[prefix + '.conv.bias'], stride=stride)
    return y[..., :y.shape[-1] - (w.shape[-1] - stride)]


def lstm(sd, prefix, x, layers):
    """x [B, C, T] -> lstm(x) + x (EncodecLSTM), written out step by step."""
    seq = x.permute(2, 0, 1)                                 # [T, B, C]
    inp = seq
    for l in range(layers):
        w_ih, w_hh = sd[f'{prefix}.lstm.weight_ih_l{l}'], sd[f'{prefix}.lstm.weight_hh_l{l}']
        bias = sd[f'{prefix}.lstm.bias_ih_l{l}'] + sd[f'{prefix}.lstm.bias_hh_l{l}']
        h = inp.new_zeros(inp.shape[1], w_hh.shape[1])
        c = torch.zeros_like(h)
        out = []
        gx = inp @ w_ih.t() + bias
        for t in range(inp.shape[0]):
            i, f, g, o = (gx[t] + h @ w_hh.t()).chunk(4, dim=-1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            out.append(h)
        inp = torch.stack(out)
    return (inp + seq).permute(1, 2, 0)


def resnet_block(sd, prefix, x):
    h = conv1d(sd, prefix + '.block.1', F.elu(x))
    h = conv1d(sd, prefix + '.block.3', F.elu(h))
    return conv1d(sd, prefix + '.shortcut', x) + h


def decode(sd, emb, cfg=None):
    """emb [B, hidden_size, T] -> waveform [B, audio_channels, T * prod(ratios)] (EncodecDecoder.forward)."""
    cfg = dict(DEFAULT, **(cfg or {}))
    x = conv1d(sd, 'layers.0', emb)
    x = lstm(sd, 'layers.1', x, cfg['num_lstm_layers'])
    i = 2
    for ratio in cfg['upsampling_ratios']:
        x = conv_transpose1d(sd, f'layers.{i + 1}', F.elu(x), ratio)
        x = resnet_block(sd, f'layers.{i + 2}', x)
        i += 3
    return conv1d(sd, f'layers.{i + 1}', F.elu(x))
