"""EnCodec (SEANet) decoder on libe2b (csrc/encodec.cu), SURVEY.md section 8f row N1.

Replaces `self.model.decoder(emb)` inside the reference's `EncodecWrapper.decode` (e2_tts_crossatt3.py:434-437; HuggingFace
`EncodecDecoder`, transformers==4.46.0 per requirements.txt:20).  Weights come from the HuggingFace decoder's state dict:
weight norm is folded once, conv weights are repacked tap-major / channels-last, every ConvTranspose1d(kernel 2s, stride s)
becomes a 2-tap causal conv with s * Co phase-major output channels, and W_hh of the LSTM is laid out per CTA.
"""
from __future__ import annotations

import torch

from . import _lib

ELU_IN, ACCUMULATE, REFLECT = 1, 2, 4
LSTM_CHUNK = 64            # sequences per LSTM launch (shared-memory footprint H * B * 4 bytes)


def _fold(sd, prefix):
    g, v = sd[prefix + '.conv.parametrizations.weight.original0'], sd[prefix + '.conv.parametrizations.weight.original1']
    return (g * v / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.ndim - 1)))).float()


class EncodecDecoderB200:
    def __init__(self, state_dict, device, upsampling_ratios=(8, 5, 4, 2), num_lstm_layers=2):
        if torch.device(device).type != 'cuda':
            raise RuntimeError('EncodecDecoderB200 runs on CUDA only: libe2b has no CPU path')
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in state_dict.items()}
        self.device = torch.device(device)
        self.ratios = tuple(upsampling_ratios)
        self.num_lstm_layers = num_lstm_layers
        self.convs = {}
        conv_names = ['layers.0']
        i = 2
        for _ in self.ratios:
            conv_names += [f'layers.{i + 2}.block.1', f'layers.{i + 2}.block.3', f'layers.{i + 2}.shortcut']
            i += 3
        conv_names.append(f'layers.{i + 1}')
        for name in conv_names:
            w = _fold(sd, name)                                            # [Co, Ci, K]
            co, ci, k = w.shape
            self.convs[name] = (w.permute(2, 1, 0).contiguous().view(k * ci, co), sd[name + '.conv.bias'].contiguous(), ci, co, k)
        self.convtrs = []
        i = 2
        for s in self.ratios:
            name = f'layers.{i + 1}'
            wt = _fold(sd, name)                                           # [Ci, Co, K = 2s]
            ci, co, k = wt.shape
            if k != 2 * s:
                raise NotImplementedError('ConvTranspose1d kernel must be twice its stride (EnCodec decoder)')
            # tap 0 multiplies x[t-1] (weights r+s), tap 1 multiplies x[t] (weights r); output channel index r * Co + co
            w2 = torch.stack((wt[:, :, s:], wt[:, :, :s]), 0)              # [2, Ci, Co, s]
            w2 = w2.permute(0, 1, 3, 2).contiguous().view(2 * ci, s * co)
            self.convtrs.append((w2, sd[name + '.conv.bias'].repeat(s).contiguous(), ci, co, s))
            i += 3
        self.lstm = []
        for l in range(num_lstm_layers):
            w_ih, w_hh = sd[f'layers.1.lstm.weight_ih_l{l}'], sd[f'layers.1.lstm.weight_hh_l{l}']
            h = w_hh.shape[1]
            bias = (sd[f'layers.1.lstm.bias_ih_l{l}'] + sd[f'layers.1.lstm.bias_hh_l{l}']).contiguous()
            whh = w_hh.view(4, h // 4, 4, h).permute(1, 3, 2, 0).contiguous()   # [cta, j, unit, gate]
            self.lstm.append((w_ih.t().contiguous(), bias, whh, h))
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)

    # ---- kernels -----------------------------------------------------------------------------------------
    def _conv(self, x, w, bias, ci, co, k, flags, y=None):
        b, t, _ = x.shape
        if y is None:
            y = torch.empty(b, t, co, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):          # launch on the tensors' device, not the process's current one
            rc = _lib.lib().e2b_conv1d_cl(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), b, t, ci, co, k, co, flags,
                                          _lib.stream_ptr(x.device))
            _lib.check(rc, None, 'e2b_conv1d_cl')
        return y

    def _named(self, name, x, flags, y=None):
        w, bias, ci, co, k = self.convs[name]
        return self._conv(x, w, bias, ci, co, k, flags, y)

    def _lstm(self, x):
        b, t, h = x.shape
        inp = x
        for l, (w_ih_t, bias, whh, hh) in enumerate(self.lstm):
            gx = self._conv(inp, w_ih_t, bias, hh, 4 * hh, 1, 0)
            out = torch.empty(b, t, hh, device=x.device, dtype=torch.float32)
            last = l == len(self.lstm) - 1
            for b0 in range(0, b, LSTM_CHUNK):
                bc = min(LSTM_CHUNK, b - b0)
                bp = (bc + 3) // 4 * 4                                     # the kernel wants a multiple of 4 sequences
                g = gx[b0:b0 + bc]
                sk = x[b0:b0 + bc] if last else None
                o = out[b0:b0 + bc]
                if bp != bc:
                    g = torch.cat((g, g.new_zeros(bp - bc, t, 4 * hh)))
                    sk = None if sk is None else torch.cat((sk, sk.new_zeros(bp - bc, t, hh)))
                    o = torch.empty(bp, t, hh, device=x.device, dtype=torch.float32)
                hbuf = torch.empty(2, hh, bp, device=x.device, dtype=torch.float32)
                with torch.cuda.device(x.device):
                    rc = _lib.lib().e2b_lstm_layer(_lib.ptr(g.contiguous()), _lib.ptr(whh), _lib.ptr(None if sk is None else sk.contiguous()), _lib.ptr(o),
                                                   _lib.ptr(hbuf), _lib.ptr(self.counter), bp, t, hh, _lib.stream_ptr(x.device))
                    _lib.check(rc, None, 'e2b_lstm_layer')
                if bp != bc:
                    out[b0:b0 + bc] = o[:bc]
            inp = out
        return inp

    @torch.no_grad()
    def __call__(self, emb):
        """emb Float[b, hidden, t] -> waveform Float[b, 1, t * prod(ratios)] (EncodecDecoder.forward)."""
        if emb.device.type != 'cuda':
            raise RuntimeError('EncodecDecoderB200 runs on CUDA tensors only')
        x = emb.to(torch.float32).transpose(1, 2).contiguous()             # channels-last
        x = self._named('layers.0', x, REFLECT)
        x = self._lstm(x)
        i = 2
        for (w2, b2, ci, co, s) in self.convtrs:
            b, t, _ = x.shape
            x = self._conv(x, w2, b2, ci, s * co, 2, ELU_IN).view(b, t * s, co)      # phase-major channels = time-major samples
            h = self._named(f'layers.{i + 2}.block.1', x, ELU_IN | REFLECT)
            y = self._named(f'layers.{i + 2}.shortcut', x, 0)
            x = self._named(f'layers.{i + 2}.block.3', h, ELU_IN | ACCUMULATE, y)
            i += 3
        x = self._named(f'layers.{i + 1}', x, ELU_IN | REFLECT)
        return x.transpose(1, 2)
