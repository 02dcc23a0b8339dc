"""B200-native drop-in for the CFM sampling path of the reference's `e2_tts_pytorch/e2_tts_crossatt3.py` ("X3").

Same names, constructor kwargs, `sample(...)` kwargs and state-dict keys as the reference classes
(`E2TTS` X3:1275, `Transformer` X3:707, `MelSpec` X3:375, `EncodecWrapper` X3:419, `DurationPredictor` X3:1147), so
`src/inference_v2a.py` / `src/inference_v2p.py` / `app.py` can import this module instead.  The modules below only
*hold parameters* (torch owns device memory, streams and checkpoints); all arithmetic of the hot path runs in
libe2b.so through the C-ABI of include/e2b.h.  There is no eager / CPU fallback: on a machine without the built
library or without a CUDA device the hot-path calls raise.

Out of scope for this path: training `forward`, duration predictor, character embeddings,
tokenizers, the frozen CLIP / T5 / Video2RollNet / EnCodec networks.  Their outputs enter `sample()` as tensors:
`text=` float [b, n, dim_text] (per-frame CLIP stream, X3:2040), `context=`/`context_mask=` (T5 output, added kwargs)
or an overridden `encode_text`, `frames=` either a precomputed piano-roll [b, n, 51] or the 5-D frame stack when a
`video2roll_net` module has been attached.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from pathlib import Path
from typing import Callable

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Module, ModuleList

from . import _lib

NOTES = 51          # X3:70


def exists(v):
    return v is not None


def default(v, d):
    return v if exists(v) else d


def lens_to_mask(t, length=None):
    """Bool[b, n] prefix mask (X3:296-305)."""
    length = int(t.amax()) if length is None else int(length)
    return torch.arange(length, device=t.device)[None, :] < t[:, None]


# --------------------------------------------------------------------------------------------------------------
# parameter holders -- names and shapes follow the reference state dict
# --------------------------------------------------------------------------------------------------------------

class _RMSNormP(Module):
    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(dim))


class _AdaNormP(Module):
    def __init__(self, dim):
        super().__init__()
        self.to_gamma = nn.Linear(dim, dim, bias=False)
        nn.init.zeros_(self.to_gamma.weight)


class _AdaLNZeroP(Module):
    def __init__(self, dim):
        super().__init__()
        self.to_gamma = nn.Linear(dim, dim)
        nn.init.zeros_(self.to_gamma.weight)
        nn.init.constant_(self.to_gamma.bias, -2.)


class _AttentionP(Module):
    def __init__(self, dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_v_head_gate = nn.Linear(dim, heads)
        nn.init.zeros_(self.to_v_head_gate.weight)
        nn.init.constant_(self.to_v_head_gate.bias, 10.)
        self.to_out = nn.Linear(inner, dim, bias=False)


class _GLUP(Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)


class _FeedForwardP(Module):
    def __init__(self, dim, mult):
        super().__init__()
        inner = int(dim * mult)
        self.ff = nn.Sequential(_GLUP(dim, inner), nn.Identity(), nn.Linear(inner, dim))


class _DepthwiseConvP(Module):
    def __init__(self, dim, kernel_size):
        super().__init__()
        self.dw_conv1d = nn.Sequential(nn.Conv1d(dim, dim, kernel_size, groups=dim, padding=kernel_size // 2), nn.SiLU())


class _CrossConditionP(Module):
    def __init__(self, dim, dim_text, dim_frames, cond_audio_to_text):
        super().__init__()
        self.text_frames_to_audio = nn.Linear(dim + dim_text + dim_frames, dim, bias=False)
        nn.init.zeros_(self.text_frames_to_audio.weight)
        if cond_audio_to_text:
            self.audio_to_text = nn.Linear(dim + dim_text, dim_text, bias=False)
            self.audio_to_frames = nn.Linear(dim + dim_frames, dim_frames, bias=False)
            nn.init.zeros_(self.audio_to_text.weight)
            nn.init.zeros_(self.audio_to_frames.weight)


class _RotaryP(Module):
    def __init__(self, dim):
        super().__init__()
        self.register_buffer('inv_freq', 1. / (10000 ** (torch.arange(0, dim, 2).float() / dim)))


class _FourierP(Module):
    def __init__(self, dim):
        super().__init__()
        self.register_buffer('weights', torch.randn(dim // 2))


def _named_tensors(module: Module, prefix: str = ''):
    out = {}
    for k, v in module.state_dict().items():
        out[prefix + k] = v
    return out


class _Engine:
    """Owns one libe2b handle for a set of weights on one device."""

    def __init__(self, cfg: dict, tensors: dict, device):
        self.lib = _lib.lib()
        if device.type != 'cuda':
            raise RuntimeError('libe2b runs on CUDA devices only (no CPU path); move the model with .to("cuda")')
        self.device = device
        self.cfg = _lib.Config(**cfg)
        self.handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.e2b_create(C.byref(self.cfg), C.byref(self.handle)), None, 'e2b_create')
            keep, arr = [], (_lib.Tensor * len(tensors))()
            for i, (name, t) in enumerate(tensors.items()):
                t = t.detach().to(device=device, dtype=torch.float32).contiguous()
                keep.append(t)
                arr[i].name = name.encode()
                arr[i].dev = t.data_ptr()
                arr[i].ndim = t.ndim
                for j, s in enumerate(t.shape):
                    arr[i].shape[j] = s
            self._check(self.lib.e2b_load_weights(self.handle, arr, len(tensors), _lib.stream_ptr(device)), 'e2b_load_weights')
            torch.cuda.synchronize(device)
        self.shape = None

    def _check(self, rc, what):
        _lib.check(rc, self.handle, what)

    def prepare(self, B, n, nc, P):
        if self.shape != (B, n, nc, P):
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                self._check(self.lib.e2b_prepare(self.handle, B, n, nc, P), 'e2b_prepare')
            self.shape = (B, n, nc, P)

    def set_conditions(self, clip, roll, ctx, lens, ctx_lens, pass_flags):
        with torch.cuda.device(self.device):
            self._check(self.lib.e2b_set_conditions(
                self.handle, _lib.ptr(clip), _lib.ptr(roll), _lib.ptr(ctx), _lib.int_array(lens), _lib.int_array(ctx_lens),
                _lib.int_array(pass_flags), _lib.stream_ptr(self.device)), 'e2b_set_conditions')

    def set_audio_cond(self, cond, cond_lens, audio_drop):
        with torch.cuda.device(self.device):
            self._check(self.lib.e2b_set_audio_cond(
                self.handle, _lib.ptr(cond), _lib.int_array(cond_lens) if cond is not None else None,
                _lib.int_array([1 if v else 0 for v in audio_drop]) if audio_drop is not None else None,
                _lib.stream_ptr(self.device)), 'e2b_set_audio_cond')

    def forward(self, x, t, P):
        B, n, d = x.shape
        pred = torch.empty(P, B, n, d, device=x.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.e2b_forward(self.handle, _lib.ptr(x), float(t), _lib.ptr(pred), _lib.stream_ptr(self.device)),
                        'e2b_forward')
        return pred

    def sample(self, y, t_grid, weights, apg, keep_parallel=0.):
        with torch.cuda.device(self.device):
            self._check(self.lib.e2b_sample(
                self.handle, _lib.ptr(y), _lib.float_array(t_grid), len(t_grid), _lib.float_array(weights) if weights else None,
                int(apg), float(keep_parallel), _lib.stream_ptr(self.device)), 'e2b_sample')
        return y

    def transformer_forward(self, x, times, lens, text, frames, ctx, ctx_lens):
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            self._check(self.lib.e2b_transformer_forward(
                self.handle, _lib.ptr(x), _lib.float_array(times), _lib.int_array(lens), _lib.ptr(text), _lib.ptr(frames),
                _lib.ptr(ctx), _lib.int_array(ctx_lens), _lib.ptr(out), _lib.stream_ptr(self.device)), 'e2b_transformer_forward')
        return out

    def flops_per_forward(self):
        return float(self.lib.e2b_forward_flops(self.handle))

    def launch_count(self):
        return int(self.lib.e2b_launch_count(self.handle))

    def __del__(self):
        try:
            if self.handle:
                self.lib.e2b_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


def _prefix_lens(mask, n, what):
    """Masks on this path are prefix masks (lens_to_mask, X3:296-305); return the per-row lengths."""
    if mask is None:
        return None
    lens = mask.sum(dim=-1)
    if not torch.equal(mask, lens_to_mask(lens, n)):
        raise NotImplementedError(f'{what}: only prefix (length) masks are supported by the CUDA path')
    return [int(v) for v in lens.tolist()]


class Transformer(Module):
    """Parameter container + CUDA forward of the reference `Transformer` (X3:707-1143)."""

    def __init__(self, *, dim, dim_text=None, dim_frames=512, depth=8, heads=8, dim_head=64, ff_mult=4, text_depth=None,
                 text_heads=None, text_dim_head=None, text_ff_mult=None, cond_on_time=True, abs_pos_emb=True, max_seq_len=8192,
                 kernel_size=31, dropout=0.1, num_registers=32,
                 attn_kwargs: dict = dict(gate_value_heads=True, softclamp_logits=True), ff_kwargs: dict = dict(),
                 if_text_modules=True, if_cross_attn=True, if_audio_conv=True, if_text_conv=False):
        super().__init__()
        assert depth % 2 == 0, 'depth needs to be even'
        dim_text = default(dim_text, dim // 2)
        unsupported = []
        if not (cond_on_time and abs_pos_emb and if_text_modules and if_cross_attn and if_audio_conv and if_text_conv):
            unsupported.append('cond_on_time/abs_pos_emb/if_text_modules/if_cross_attn/if_audio_conv/if_text_conv must all be True')
        if default(text_depth, depth) != depth or default(text_heads, heads) != heads or default(text_dim_head, dim_head) != dim_head \
                or default(text_ff_mult, ff_mult) != ff_mult:
            unsupported.append('text_depth/text_heads/text_dim_head/text_ff_mult must equal the audio values')
        if dim_head != 64 or kernel_size != 31 or ff_kwargs or dict(attn_kwargs) != dict(gate_value_heads=True, softclamp_logits=True):
            unsupported.append('dim_head=64, kernel_size=31, default attn_kwargs/ff_kwargs only')
        if unsupported:
            raise NotImplementedError('libe2b implements the shipped architecture family (src/inference_v2a.py:76-90): ' + '; '.join(unsupported))

        self.dim, self.dim_text, self.dim_frames = dim, dim_text, dim_frames
        self.depth, self.heads, self.dim_head, self.ff_mult = depth, heads, dim_head, ff_mult
        self.max_seq_len, self.num_registers, self.kernel_size = max_seq_len, num_registers, kernel_size
        self.cond_on_time, self.if_cross_attn, self.if_audio_conv, self.if_text_conv = True, True, True, True
        self.frames_heads = 8                                            # hard-coded at X3:914

        self.abs_pos_emb = nn.Embedding(max_seq_len, dim)
        self.registers = nn.Parameter(torch.zeros(num_registers, dim))
        self.text_registers = nn.Parameter(torch.zeros(num_registers, dim_text))
        self.frames_registers = nn.Parameter(torch.zeros(num_registers, dim_frames))
        for p in (self.registers, self.text_registers, self.frames_registers):
            nn.init.normal_(p, std=0.02)
        self.rotary_emb, self.text_rotary_emb, self.frames_rotary_emb = _RotaryP(dim_head), _RotaryP(dim_head), _RotaryP(dim_head)
        self.time_cond_mlp = nn.Sequential(_FourierP(dim), nn.Linear(dim + 1, dim), nn.SiLU())

        self.layers = ModuleList([])
        for ind in range(depth):
            speech = ModuleList([
                nn.Linear(dim * 2, dim, bias=False) if ind >= depth // 2 else None,
                _DepthwiseConvP(dim, kernel_size),
                _AdaNormP(dim), _AttentionP(dim, heads, dim_head), _AdaLNZeroP(dim),
                _AdaNormP(dim), _AttentionP(dim, heads, dim_head), _AdaLNZeroP(dim),
                _AdaNormP(dim), _FeedForwardP(dim, ff_mult), _AdaLNZeroP(dim),
            ])
            text = ModuleList([
                _DepthwiseConvP(dim_text, kernel_size),
                _RMSNormP(dim_text), _AttentionP(dim_text, heads, dim_head),
                _RMSNormP(dim_text), _FeedForwardP(dim_text, ff_mult),
                _CrossConditionP(dim, dim_text, dim_frames, cond_audio_to_text=ind != depth - 1),
            ])
            frames = ModuleList([
                _DepthwiseConvP(dim_frames, kernel_size),
                _RMSNormP(dim_frames), _AttentionP(dim_frames, self.frames_heads, 64),
                _RMSNormP(dim_frames), _FeedForwardP(dim_frames, 4),
            ])
            self.layers.append(ModuleList([speech, text, frames]))
        self.final_norm = _RMSNormP(dim)
        self._engine = None

    # ---- engine management -------------------------------------------------------------------------------
    precision = 'bf16'      # 'bf16' (tensor-core path, rel-L2 <= 1e-2) or 'fp32' (error-compensated mode, <= 1e-4)

    def engine_config(self, num_channels=64, precision=None):
        return dict(precision=_lib.PRECISIONS[precision or self.precision], depth=self.depth, dim=self.dim, dim_text=self.dim_text, dim_frames=self.dim_frames, heads=self.heads,
                    dim_head=self.dim_head, frames_heads=self.frames_heads, num_channels=num_channels,
                    num_registers=self.num_registers, kernel_size=self.kernel_size, notes=NOTES, max_seq_len=self.max_seq_len,
                    ff_mult=self.ff_mult)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def _standalone_engine(self):
        fp = (self.precision, _weights_fingerprint(self))
        if self._engine is not None and self.__dict__.get('_engine_fp') != fp:
            self._engine = None                      # precision switch or weights edited since they were packed
        self.__dict__['_engine_fp'] = fp
        if self._engine is None:
            dev = self.registers.device
            tensors = _named_tensors(self, 'transformer.')
            z = lambda *s: torch.zeros(*s, device=dev)
            tensors.update({'proj_in.weight': z(self.dim, 64), 'proj_in.bias': z(self.dim), 'to_pred.weight': z(64, self.dim),
                            'to_pred.bias': z(64), 'proj_frames.weight': z(self.dim_frames, NOTES), 'proj_frames.bias': z(self.dim_frames)})
            self._engine = _Engine(self.engine_config(64), tensors, dev)
            self._engine_precision = self.precision
        return self._engine

    @torch.no_grad()
    def forward(self, x, times=None, mask=None, text_embed=None, frames_embed=None, context=None, context_mask=None):
        """Float[b n d] -> Float[b n d]; same arguments as the reference (X3:941-950)."""
        assert exists(times), '`times` must be passed in if `cond_on_time` is set to `True` and vice versa'   # X3:953
        b, n, _ = x.shape
        assert n <= self.max_seq_len, f'{n} exceeds the set `max_seq_len` ({self.max_seq_len}) on Transformer'  # X3:958
        if text_embed is None or frames_embed is None or context is None:
            raise NotImplementedError('the CUDA path needs text_embed, frames_embed and context (the shipped call always passes them)')
        if x.shape[-1] != self.dim or tuple(text_embed.shape) != (b, n, self.dim_text) or tuple(frames_embed.shape) != (b, n, self.dim_frames) \
                or context.ndim != 3 or context.shape[0] != b or context.shape[-1] != self.dim:
            raise ValueError(f'Transformer.forward: x {tuple(x.shape)}, text_embed {tuple(text_embed.shape)}, frames_embed '
                             f'{tuple(frames_embed.shape)}, context {tuple(context.shape)} do not fit dim {self.dim} / dim_text '
                             f'{self.dim_text} / dim_frames {self.dim_frames}')
        eng = self._standalone_engine()
        if times.ndim == 0:
            times = times.expand(b)
        eng.prepare(b, n, context.shape[1], 1)
        f32 = lambda t: t.to(dtype=torch.float32).contiguous()
        lens = _prefix_lens(mask, n, 'mask') or [n] * b
        ctx_lens = _prefix_lens(context_mask, context.shape[1], 'context_mask') or [context.shape[1]] * b
        return eng.transformer_forward(f32(x), [float(v) for v in times.tolist()], lens, f32(text_embed), f32(frames_embed),
                                       f32(context), ctx_lens)


class DurationPredictor(Module):
    """Out of scope (disabled in every shipped config: `duration_predictor=None`, src/inference_v2a.py:72)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError('DurationPredictor is outside the CFM sampling path (SURVEY.md section 8f, row N4)')


class MelSpec(Module):
    """STFT + mel front end on the GPU (reference MelSpec X3:375-417: torchaudio MelSpectrogram, power 1, HTK, norm None,
    center/reflect, then log(clamp(1e-5)))."""

    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=100, sampling_rate=24_000,
                 normalize=False, power=1, norm=None, center=True):
        super().__init__()
        if normalize or power != 1 or norm is not None or not center or win_length != filter_length:
            raise NotImplementedError('libe2b MelSpec implements the reference defaults (power=1, norm=None, center=True, win=n_fft)')
        self.n_fft, self.hop, self.n_mel_channels, self.sampling_rate = filter_length, hop_length, n_mel_channels, sampling_rate
        self.register_buffer('window', torch.hann_window(win_length, periodic=True), persistent=False)
        self.register_buffer('fb', self.mel_filterbank(filter_length // 2 + 1, 0., sampling_rate / 2., n_mel_channels, sampling_rate),
                             persistent=False)

    @staticmethod
    def mel_filterbank(n_freqs, f_min, f_max, n_mels, sample_rate):
        """HTK triangular filters, no area normalisation: [n_freqs, n_mels]."""
        hz2mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)
        all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
        m_pts = torch.linspace(hz2mel(f_min), hz2mel(f_max), n_mels + 2)
        f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
        down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
        up = slopes[:, 2:] / f_diff[1:]
        return torch.clamp(torch.min(down, up), min=0.).contiguous()

    @torch.no_grad()
    def forward(self, inp):
        if inp.ndim == 3:
            inp = inp.squeeze(1)                       # 'b 1 nw -> b nw'
        assert inp.ndim == 2
        if inp.device.type != 'cuda':
            raise RuntimeError('libe2b MelSpec runs on CUDA tensors only (no CPU path)')
        if self.window.device != inp.device:
            self.to(inp.device)
        wav = inp.to(torch.float32).contiguous()
        B, nw = wav.shape
        T = nw // self.hop + 1
        out = torch.empty(B, self.n_mel_channels, T, device=wav.device, dtype=torch.float32)
        with torch.cuda.device(wav.device):
            _lib.check(_lib.lib().e2b_melspec(_lib.ptr(wav), B, nw, self.n_fft, self.hop, self.n_mel_channels, _lib.ptr(self.window),
                                              _lib.ptr(self.fb), _lib.ptr(out), _lib.stream_ptr(wav.device)), None, 'e2b_melspec')
        return out


class EncodecWrapper(Module):
    """EnCodec latent encoder / decoder wrapper (reference X3:419-437).  The weights and the encoder stay HuggingFace
    `transformers` like in the reference; the decoder (the vocoder downstream of the sampler) runs
    on libe2b."""

    def __init__(self, path):
        super().__init__()
        from transformers import AutoProcessor, EncodecModel
        self.model = EncodecModel.from_pretrained(path)
        self.processor = AutoProcessor.from_pretrained(path)
        for p in self.model.parameters():
            p.requires_grad = False
        self.model.eval()

    def forward(self, waveform):
        with torch.no_grad():
            inputs = self.processor(raw_audio=waveform[0], sampling_rate=self.processor.sampling_rate, return_tensors='pt')
            return self.model.encoder(inputs.input_values)

    def decode_batch(self, emb):
        """Float[b, 128, t] latents -> Float[b, 1, 320 t] waveforms on libe2b (csrc/encodec.cu)."""
        if not emb.is_cuda:
            raise RuntimeError('EncodecWrapper.decode runs on CUDA tensors only: libe2b has no CPU path')
        dec = self.__dict__.get('_b200')
        if dec is None or dec.device != emb.device:
            from .encodec import EncodecDecoderB200
            cfg = self.model.config
            dec = EncodecDecoderB200(self.model.decoder.state_dict(), emb.device, upsampling_ratios=tuple(cfg.upsampling_ratios),
                                     num_lstm_layers=cfg.num_lstm_layers)
            self.__dict__['_b200'] = dec
        return dec(emb)

    def decode(self, emb):
        """Reference semantics (X3:434-437): the decoder output of the FIRST batch item, Float[1, samples]."""
        return self.decode_batch(emb)[0]


# 'null' is the reference's CFG pass (X3:2104: drop_audio_cond, drop_text_cond, drop_text_prompt); the others compose its
# public drop flags one at a time
GUIDANCE_PASSES = {'null': _lib.DROP_CLIP | _lib.DROP_CTX | _lib.DROP_AUDIO, 'drop_t5': _lib.DROP_CTX, 'drop_clip': _lib.DROP_CLIP,
                   'drop_roll': _lib.DROP_ROLL, 'drop_audio': _lib.DROP_AUDIO}


def _weights_fingerprint(module):
    """Cheap identity of a module's tensors: storage address + in-place version counter of every parameter / buffer."""
    return tuple((k, v.data_ptr(), v._version, str(v.device)) for k, v in module.state_dict(keep_vars=True).items())


class E2TTS(Module):
    """CFM wrapper with the reference constructor surface (X3:1278-1318) and `sample` (X3:2127-2305)."""

    def __init__(self, transformer: dict | Transformer = None, duration_predictor=None,
                 odeint_kwargs: dict = dict(method='euler'), audiocond_drop_prob=0.30, cond_drop_prob=0.20, prompt_drop_prob=0.10,
                 num_channels=None, mel_spec_module: Module | None = None, char_embed_kwargs: dict = dict(),
                 mel_spec_kwargs: dict = dict(), frac_lengths_mask=(0.7, 1.), audiocond_snr=None, concat_cond=False,
                 interpolated_text=False, text_num_embeds=None, tokenizer='char_utf8', use_vocos=True,
                 pretrained_vocos_path='charactr/vocos-mel-24khz', sampling_rate=None, frame_size: int = 320,
                 velocity_consistency_weight=-1e-5, if_cond_proj_in=True, cond_proj_in_bias=True, if_embed_text=True,
                 if_text_encoder2=True, if_clip_encoder=False, video_encoder='clip_vit'):
        super().__init__()
        if isinstance(transformer, dict):
            transformer = Transformer(**transformer, cond_on_time=True)
        if duration_predictor is not None:
            raise NotImplementedError('duration_predictor is outside the CFM sampling path (shipped configs pass None)')
        if dict(odeint_kwargs).get('method', 'euler') != 'euler':
            raise NotImplementedError('only the fixed-grid Euler solver the shipped configs use is implemented')
        if concat_cond:
            raise NotImplementedError('concat_cond is not used by the shipped configs')
        self.transformer = transformer
        dim, dim_text, dim_frames = transformer.dim, transformer.dim_text, transformer.dim_frames
        self.dim, self.dim_text = dim, dim_text
        self.frac_lengths_mask, self.audiocond_snr = frac_lengths_mask, audiocond_snr
        self.duration_predictor = None
        self.odeint_kwargs = odeint_kwargs
        self.mel_spec = default(mel_spec_module, None)
        self.num_channels = num_channels
        self.sampling_rate = default(sampling_rate, None)
        self.frame_size = frame_size
        self.concat_cond = False
        self.proj_in = nn.Linear(num_channels, dim)
        self.cond_proj_in = nn.Linear(num_channels, dim, bias=cond_proj_in_bias) if if_cond_proj_in else None
        self.to_pred = nn.Linear(dim, num_channels)
        self.audiocond_drop_prob, self.cond_drop_prob, self.prompt_drop_prob = audiocond_drop_prob, cond_drop_prob, prompt_drop_prob
        self.tokenizer = tokenizer
        self.embed_text = None
        if if_embed_text:
            raise NotImplementedError('character text embedding (if_embed_text) is disabled in the shipped configs and not on this path')
        self.register_buffer('zero', torch.tensor(0.), persistent=False)
        self.velocity_consistency_weight = velocity_consistency_weight
        self.vocos = None
        if if_text_encoder2:        # frozen FLAN-T5 prompt encoder (X3:1411-1416); third-party, stays HuggingFace
            from transformers import AutoTokenizer, T5EncoderModel
            self.tokenizer2 = AutoTokenizer.from_pretrained('./ckpts/flan-t5-large')
            self.text_encoder2 = T5EncoderModel.from_pretrained('./ckpts/flan-t5-large')
            for p in self.text_encoder2.parameters():
                p.requires_grad = False
            self.text_encoder2.eval()
        self.proj_text = None
        self.proj_frames = nn.Linear(NOTES, dim_frames)
        self.image_processor, self.image_encoder = None, None
        if if_clip_encoder:         # frozen CLIP image encoder (X3:1420-1425); third-party, stays HuggingFace
            if video_encoder != 'clip_vit':
                raise NotImplementedError('only video_encoder="clip_vit" (the shipped value) is supported')
            from transformers import CLIPImageProcessor, CLIPVisionModelWithProjection
            self.image_processor = CLIPImageProcessor()
            self.image_encoder = CLIPVisionModelWithProjection.from_pretrained('./ckpts/IP-Adapter/', subfolder='sdxl_models/image_encoder')
            for p in self.image_encoder.parameters():
                p.requires_grad = False
            self.image_encoder.eval()
        self.video_encoder = video_encoder
        self.video2roll_net = None  # attach the Audeo Video2RollNet (src/audeo/Video2RollNet.py) here to feed raw frames
        self._engine = None
        self._last_launches = 0

    # ---- plumbing ----------------------------------------------------------------------------------------
    @property
    def device(self):
        return next(self.parameters()).device

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **k):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **k)

    precision = 'bf16'      # set to 'fp32' for the error-compensated mode (rel-L2 <= 1e-4 vs the fp32 reference)

    def engine(self) -> _Engine:
        """Repack the current parameters into libe2b (done once per load / device move / precision change)."""
        fp = (self.precision, _weights_fingerprint(self))
        if self._engine is not None and self.__dict__.get('_engine_fp') != fp:
            # precision switch, load_state_dict on a sub-module, EMA copy or any in-place edit (p.data.copy_, optimiser step):
            # the packed bf16 copy inside libe2b is stale
            self._engine = None
        if self._engine is None:
            tensors = _named_tensors(self.transformer, 'transformer.')
            for name in ('proj_in', 'to_pred', 'proj_frames', 'cond_proj_in'):
                m = getattr(self, name)
                if m is None:
                    continue
                tensors[name + '.weight'] = m.weight
                if m.bias is not None:
                    tensors[name + '.bias'] = m.bias
            self._engine = _Engine(self.transformer.engine_config(self.num_channels, self.precision), tensors, self.device)
            self._engine_precision = self.precision
            self.__dict__['_engine_fp'] = fp
        return self._engine

    # ---- condition encoders (outside the hot path) -------------------------------------------------------
    def encode_text(self, prompt):
        """FLAN-T5 encoder output and boolean mask (X3:1648-1657)."""
        if not hasattr(self, 'text_encoder2'):
            raise RuntimeError('no T5 encoder attached: pass context=/context_mask= to sample() or override encode_text')
        device = self.device
        batch = self.tokenizer2(prompt, max_length=self.tokenizer2.model_max_length, padding=True, truncation=True, return_tensors='pt')
        ids, am = batch.input_ids.to(device), batch.attention_mask.to(device)
        with torch.no_grad():
            hidden = self.text_encoder2(input_ids=ids, attention_mask=am)[0]
        return hidden, (am == 1)

    CONTEXT_CACHE_SIZE = 256

    def encode_text_cached(self, prompt):
        """T5 context for a batch of prompts, encoding every distinct prompt once and remembering it across calls (SURVEY.md 8f
        N2: the reference re-encodes the whole batch on every network call, X3:2057, i.e. 2 x steps times per sample()).  The
        per-prompt rows are the encoder output over the prompt's own tokens; padded positions are zero and masked, exactly what
        the cross-attention key mask (X3:1131) makes of the reference's padded batch."""
        cache = self.__dict__.setdefault('_ctx_cache', {})
        prompt = list(prompt)
        missing = [p for p in dict.fromkeys(prompt) if p not in cache]
        if missing:
            hidden, mask = self.encode_text(missing)
            for i, p in enumerate(missing):
                n = int(mask[i].sum())
                if not bool(mask[i, :n].all()):
                    raise RuntimeError('encode_text returned a mask that is not a prefix (right padding expected)')
                cache[p] = hidden[i, :n].detach().to(torch.float32).clone()
            while len(cache) > self.CONTEXT_CACHE_SIZE:
                cache.pop(next(iter(cache)))
        rows = [cache[p] for p in prompt]
        nc = max(int(r.shape[0]) for r in rows)
        context = rows[0].new_zeros(len(rows), nc, rows[0].shape[1])
        context_mask = torch.zeros(len(rows), nc, dtype=torch.bool, device=rows[0].device)
        for i, r in enumerate(rows):
            context[i, :r.shape[0]] = r
            context_mask[i, :r.shape[0]] = True
        return context, context_mask

    VIDEO_FEATURE_SUFFIX = {'clip_vit': '.generated.npz', 'clip_vit2': '.generated.clip_vit2.npz', 'clip_convnext': '.generated.clip_convnext.npz',
                            'dinov2': '.generated.dinov2.npz', 'mixed': '.generated.mixed.npz'}
    # videos under the training set's root keep their features in a sibling directory instead (X3:1679-1691); the root is the
    # reference author's path -- point it at a local copy of the dataset to reuse caches the reference wrote there
    VGGSOUND_ROOT = '/ailab-train2/speech/zhanghaomin/VGGSound/'
    VGGSOUND_FEATURE_DIR = {'clip_vit': '/feature/', 'clip_vit2': '/feature_clip_vit2/', 'clip_convnext': '/feature_clip_convnext/',
                            'dinov2': '/feature_dinov2/', 'mixed': '/feature_mixed/'}

    def video_feature_path(self, video_path):
        """Cache file of a clip's per-video-frame embeddings (X3:1679-1704)."""
        if self.video_encoder not in self.VIDEO_FEATURE_SUFFIX:
            raise Exception('Invalid video_encoder ' + str(self.video_encoder))
        if video_path.startswith(self.VGGSOUND_ROOT):
            return video_path.replace('/video/', self.VGGSOUND_FEATURE_DIR[self.video_encoder]).replace('.mp4', '.npz')
        return video_path.replace('.mp4', self.VIDEO_FEATURE_SUFFIX[self.video_encoder])

    def encode_video(self, video_paths, l):
        """Per-latent-frame CLIP stream Float[b, l, d] from the reference's feature caches (X3:1659-1827).

        Each entry of `video_paths` is None (all-zero clip), a path, or (path, start_sample, max_sample).  The cache is the
        `.npz` the reference writes next to the video (arr_0 = embeddings [F, d] fp32, arr_1 = duration in seconds, X3:1796);
        running the CLIP image encoder itself is out of scope, so a missing cache is an error here.  The nearest-video-frame
        resampling and the zero padding run on the GPU (csrc/staging.cu)."""
        d = self.dim_text if self.proj_text is None else self.dim_text_raw
        device = self.device
        if torch.device(device).type != 'cuda':
            raise RuntimeError('E2TTS.encode_video stages on CUDA only: libe2b has no CPU path')
        embs, meta, durs = [], [], []
        off = 0
        for video_path in video_paths:
            if video_path is None:                                                   # X3:1670-1673
                meta.append((0, 2, 0, 0))
                durs.append(1.0)
                continue
            if isinstance(video_path, tuple):                                        # X3:1675-1678
                video_path, start_sample, max_sample = video_path
            else:
                start_sample, max_sample = 0, None
            feature_path = self.video_feature_path(video_path)
            if not os.path.exists(feature_path):
                raise RuntimeError(f'no cached video features at {feature_path}: extracting them (CLIP image encoder) is outside the '
                                   'hot path -- run the reference once to create the cache, or pass text=Float[b, n, dim_text]')
            data = np.load(feature_path)
            emb = np.ascontiguousarray(data['arr_0'], dtype=np.float32)
            duration = float(data['arr_1'].item())
            if emb.ndim != 2 or emb.shape[1] != d:
                raise RuntimeError(f'{feature_path}: expected embeddings [F, {d}], got {tuple(emb.shape)}')
            if emb.shape[0] < 2:
                raise ZeroDivisionError('float division by zero')                    # duration / (F - 1), X3:1806
            if max_sample is None:
                max_sample = int(duration * self.sampling_rate)                      # X3:1802-1803
            count = min(l, len(range(start_sample, max_sample, self.frame_size)))     # X3:1805-1810
            if count == 0:
                raise RuntimeError('torch.cat(): expected a non-empty list of Tensors')   # what X3:1811 raises
            embs.append(torch.from_numpy(emb))
            meta.append((off, emb.shape[0], count, start_sample))
            durs.append(duration)
            off += emb.shape[0]
        out = torch.empty(len(video_paths), l, d, device=device, dtype=torch.float32)
        emb_dev = (torch.cat(embs, 0) if embs else torch.zeros(2, d)).to(device, non_blocking=True)
        meta_dev = torch.tensor(meta, dtype=torch.int64).to(device, non_blocking=True)
        dur_dev = torch.tensor(durs, dtype=torch.float64).to(device, non_blocking=True)
        with torch.cuda.device(device):
            rc = _lib.lib().e2b_stage_clip(_lib.ptr(emb_dev), _lib.ptr(meta_dev), _lib.ptr(dur_dev), len(video_paths), l, d,
                                           int(self.sampling_rate), int(self.frame_size), _lib.ptr(out), _lib.stream_ptr(device))
            _lib.check(rc, None, 'e2b_stage_clip')
        return out

    def encode_frames(self, x, l):
        """Piano-roll stream Float[b, l, 51] from the grey-scale frame stack Float[b, 1, t, w, h] (X3:1525-1555; SURVEY.md 8f
        N3): 5-frame clamped windows (csrc/frames.cu, replaces the reference's Python loop over t) -> the attached Video2RollNet
        (third-party ResNet18+FPN, stays torch) -> sigmoid, x3 repeat, cut / zero-pad to l frames (one kernel)."""
        if self.video2roll_net is None:
            raise NotImplementedError('attach `video2roll_net` or pass frames as a precomputed roll Float[b, n, 51]')
        b, c, t, w, h = x.shape
        assert c == 1
        if x.device.type != 'cuda':
            raise RuntimeError('E2TTS.encode_frames runs on CUDA tensors only: libe2b has no CPU path')
        L = _lib.lib()
        x = x.to(torch.float32).contiguous()
        win = torch.empty(b * t, 5, w, h, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(L.e2b_frame_windows(_lib.ptr(x), _lib.ptr(win), b, t, w * h, 5, _lib.stream_ptr(x.device)), None, 'e2b_frame_windows')
            logits = self.video2roll_net(win).to(torch.float32).contiguous()             # [b*t, 51]
            assert tuple(logits.shape) == (b * t, NOTES), logits.shape
            roll = torch.empty(b, l, NOTES, device=x.device, dtype=torch.float32)
            _lib.check(L.e2b_roll_expand(_lib.ptr(logits), _lib.ptr(roll), b, t, l, NOTES, 3, _lib.stream_ptr(x.device)), None, 'e2b_roll_expand')
        return roll

    FRAMES_RAW_SUFFIX = '.generated_frames_raw.2.npz'      # X3:1886
    VIDEO_MULTI = 3.0                                        # latent frames per roll frame, X3:1931

    @staticmethod
    def encode_video_frames(video_paths, l, piano=True):
        """Grey-scale frame stacks for the piano-roll net, from the reference's caches (static, host side; X3:1829-1991 --
        `app.py:236` / `predict.py` call it as `E2TTS.encode_video_frames(video_paths, l, piano)`; the dataset collate calls it with
        two arguments, trainer3:1377, hence the default).

        Returns `(video_frames Float[b', 1, T, 100, 900], midis Float[b', l, 51])` on the CPU like the reference, or
        `(None, None)` when no clip contributes (no piano clip, X3:1961-1962).  Clips that are None or not piano are *dropped*
        from the batch, as in the reference.  The cache is the `.generated_frames_raw.2.npz` the reference writes next to the
        video (arr_0 = frames [F, 100, 900, 1] fp32, arr_1 = duration in seconds, X3:1905); decoding the video itself (moviepy +
        PIL) is outside the hot path, so a missing cache is an error here.  Frame k of the result is video frame
        j = min(round(i / 24000 / (duration / F)), F - 1) for i = start, start + 960, ... (X3:1934-1940)."""
        video_frames, video_lens = [], []
        fsv = int(E2TTS.VIDEO_MULTI * 320)
        want = math.floor(l / E2TTS.VIDEO_MULTI) + 1
        for video_path in video_paths:
            if video_path is None or not piano:                                      # X3:1872-1876, 1924-1928
                video_lens.append(0)
                continue
            if isinstance(video_path, tuple):
                video_path, start_sample, max_sample = video_path
            else:
                start_sample, max_sample = 0, None
            raw_path = video_path.replace('.mp4', E2TTS.FRAMES_RAW_SUFFIX)
            if not os.path.exists(raw_path):
                raise RuntimeError(f'no cached frames at {raw_path}: decoding the video (moviepy / PIL) is outside the hot path -- run the '
                                   'reference once to create the cache, or pass frames= / a precomputed roll to sample()')
            data = np.load(raw_path)
            frames_raw = torch.from_numpy(np.ascontiguousarray(data['arr_0'], dtype=np.float32))
            duration = float(data['arr_1'].item())
            F_ = frames_raw.shape[0]
            if max_sample is None:
                max_sample = int(duration * 24000)
            idx = []
            for i in range(start_sample, max_sample + fsv, fsv):
                idx.append(min(round(i / 24000 / (duration / (F_ - 0))), F_ - 1))
                if len(idx) >= want:
                    break
            picked = frames_raw[torch.tensor(idx, dtype=torch.long)]
            video_frames.append(picked.unsqueeze(0))
            video_lens.append(picked.shape[0])
        if len(video_frames) == 0:
            return None, None
        t_max = max(want, max(video_lens))
        for i, vf in enumerate(video_frames):
            if vf.shape[1] < t_max:
                video_frames[i] = torch.cat([vf, torch.zeros(1, t_max - vf.shape[1], *vf.shape[2:])], 1)
        video_frames = torch.cat(video_frames, 0).permute(0, 4, 1, 2, 3)
        midis = torch.zeros(video_frames.shape[0], l, NOTES)
        return video_frames, midis

    # ---- sampling ----------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, cond, *, text=None, lens=None, duration=None, steps=32, cfg_strength=1., remove_parallel_component=True,
               sway_sampling=True, max_duration=4096, vocoder: Callable | None = None, return_raw_output=None,
               save_to_filename=None, prompt=None, video_drop_prompt=None, audio_drop_prompt=None, video_paths=None,
               frames=None, midis=None, context=None, context_mask=None, guidance=None, noise=None, keep_parallel_frac=0.):
        """Reference signature (X3:2128-2148) plus `context`/`context_mask` (precomputed T5 output), `guidance`
        (K-pass generalisation: list of (pass_name, weight)) and `noise` (explicit y0)."""
        self.eval()
        if cond.ndim == 2:                                                           # raw wave, X3:2157-2160
            cond = self.mel_spec(cond).transpose(1, 2)
            assert cond.shape[-1] == self.num_channels
        batch, cond_seq_len, device = *cond.shape[:2], cond.device
        if device.type != 'cuda':
            raise RuntimeError('E2TTS.sample runs on CUDA tensors only: libe2b has no CPU path')

        if frames is None:                                                           # X3:2164-2176
            roll = None
        elif frames.ndim == 3:
            roll = frames
        else:
            roll = self.encode_frames(frames, cond_seq_len)
        if roll is not None:
            if roll.shape[1] < cond_seq_len:
                roll = torch.cat([roll, torch.zeros(batch, cond_seq_len - roll.shape[1], NOTES, device=device)], 1)
            roll = roll[:, :cond_seq_len].to(device=device, dtype=torch.float32).contiguous()

        if not exists(lens):
            lens = torch.full((batch,), cond_seq_len, device=device, dtype=torch.long)
        if video_paths is not None:
            text = self.encode_video(video_paths, cond_seq_len)
        if not (torch.is_tensor(text) and text.ndim == 3):
            raise NotImplementedError('pass the CLIP stream as text=Float[b, n, dim_text] (character text is not on this path)')

        if exists(duration):                                                         # X3:2198-2206
            if isinstance(duration, int):
                duration = torch.full((batch,), duration, device=device, dtype=torch.long)
        else:
            duration = lens
        lens = lens.to(device)
        duration = torch.maximum(lens, duration.to(device)).clamp(max=max_duration)
        assert duration.shape[0] == batch
        n = int(duration.amax())
        # X3:2224 / 2259: ONE switch for the whole batch, taken from item 0 -- lens < duration means audio-conditioned in-painting
        inpaint = int(lens[0]) != int(duration[0])
        if inpaint:
            if self.cond_proj_in is None:
                raise TypeError("'NoneType' object is not callable: in-painting (lens < duration) needs E2TTS(if_cond_proj_in=True), "
                                "the reference fails the same way at X3:2034")
            if self.audiocond_snr is not None:
                raise IndexError('audiocond_snr is set: the reference raises here too (add_noise indexes the [b, n, d] condition with the '
                                 '[b, n, 1] cond_mask, X3:2123/2214); sample with audiocond_snr=None')

        # shapes the engine will read through raw pointers (the reference would raise a shape error from the first matmul / cat)
        if text.shape[0] != batch or text.shape[-1] != self.dim_text:
            raise ValueError(f'text must be Float[{batch}, n, {self.dim_text}], got {tuple(text.shape)}')
        if cond.shape[-1] != self.num_channels:
            raise ValueError(f'cond must be Float[b, n, {self.num_channels}], got {tuple(cond.shape)}')
        if roll is not None and (roll.shape[0] != batch or roll.shape[-1] != NOTES):
            raise ValueError(f'frames (piano roll) must be Float[{batch}, n, {NOTES}], got {tuple(roll.shape)}')
        if text.shape[1] < min(n, cond_seq_len):
            raise ValueError(f'text has {text.shape[1]} frames, the latents {cond_seq_len}')
        if n > text.shape[1]:
            text = F.pad(text, (0, 0, 0, n - text.shape[1]))
        if roll is not None and n > roll.shape[1]:
            roll = F.pad(roll, (0, 0, 0, n - roll.shape[1]))
        clip = text[:, :n].to(device=device, dtype=torch.float32).contiguous()
        roll = None if roll is None else roll[:, :n].contiguous()

        # T5 context (hoisted: the reference re-encodes it on every network call, X3:2057)
        if context is None:
            if prompt is None:
                raise NotImplementedError('a prompt (or context=) is required: the shipped callers always pass one')
            prompt = list(prompt)
            if len(prompt) != batch:
                raise ValueError(f'{len(prompt)} prompts for a batch of {batch}')
            if video_drop_prompt is not None:
                for b in range(batch):
                    if video_drop_prompt[b]:
                        prompt[b] = 'the sound of X X'                               # X3:2053-2056
            context, context_mask = self.encode_text_cached(prompt)
        if context.ndim != 3 or context.shape[0] != batch or context.shape[-1] != self.dim:
            raise ValueError(f'context must be Float[{batch}, nc, {self.dim}], got {tuple(context.shape)}')
        if context_mask is not None and tuple(context_mask.shape) != tuple(context.shape[:2]):
            raise ValueError(f'context_mask must be Bool{list(context.shape[:2])}, got {tuple(context_mask.shape)}')
        context = context.to(device=device, dtype=torch.float32).clone()
        if video_drop_prompt is not None:
            for b in range(batch):
                if video_drop_prompt[b]:
                    context[b] = 0                                                   # X3:2061-2062
        nc = context.shape[1]
        ctx_lens = _prefix_lens(context_mask, nc, 'context_mask') if context_mask is not None else [nc] * batch

        # guidance passes: plain CFG (X3:2090-2113) is [('null', cfg_strength)]
        if guidance is None:
            guidance = [('null', float(cfg_strength))] if cfg_strength >= 1e-5 else []
        guidance = [(k, float(w)) for k, w in guidance if abs(w) >= 1e-5]
        guidance.sort(key=lambda kw: (GUIDANCE_PASSES[kw[0]] & _lib.DROP_CTX) != 0)   # context-live passes first
        apg = bool(remove_parallel_component) and len(guidance) > 0
        if apg and len(guidance) != 1:
            raise NotImplementedError('remove_parallel_component is defined for plain 2-pass CFG only')
        flags = [0] + [GUIDANCE_PASSES[k] for k, _ in guidance]
        weights = [w for _, w in guidance]

        cond = F.pad(cond, (0, 0, 0, n - cond_seq_len), value=0.)                    # X3:2212
        if noise is not None and tuple(noise.shape) != tuple(cond.shape):
            raise ValueError(f'noise must have the shape of the (padded) latents {tuple(cond.shape)}, got {tuple(noise.shape)}')

        eng = self.engine()
        eng.prepare(batch, n, nc, len(flags))
        eng.set_conditions(clip, roll, context.contiguous(), [int(v) for v in duration.tolist()], ctx_lens, flags)
        cond_f = None
        if inpaint:
            # step_cond = where(cond_mask, cond, 0) enters every network call through cond_proj_in (X3:2228, 2029-2035; the null pass
            # and clips flagged by audio_drop_prompt see zeros, X3:2016-2020); cond_mask = lens_to_mask(lens) (X3:2196, 2213)
            cond_f = cond.to(torch.float32).contiguous()
            eng.set_audio_cond(cond_f, [int(v) for v in lens.tolist()], audio_drop_prompt)

        y = torch.randn_like(cond) if noise is None else noise.to(device=device, dtype=cond.dtype).clone()   # X3:2248
        y = y.to(torch.float32).contiguous()
        t = torch.linspace(0, 1, steps, device=self.device)                          # X3:2250-2252
        if sway_sampling:
            t = t + -1.0 * (torch.cos(torch.pi / 2 * t) - 1 + t)
        before = eng.launch_count()
        out = eng.sample(y, [float(v) for v in t.tolist()], weights, apg, keep_parallel_frac)
        self._last_launches = eng.launch_count() - before
        if inpaint and steps < 2:                                                    # no Euler update to fold the select into
            out = torch.where(lens_to_mask(lens, n)[..., None], cond_f, out)         # X3:2259-2260

        if exists(return_raw_output) and return_raw_output:                          # X3:2265-2266
            return out
        mask = lens_to_mask(duration.to(device), n)
        if exists(vocoder):                                                          # X3:2270-2287
            assert not exists(self.vocos), '`use_vocos` should not be turned on if you are passing in a custom `vocoder` on sampling'
            out = vocoder(out.transpose(1, 2))
        elif exists(self.vocos):
            if hasattr(self.vocos, 'decode_batch') and int(duration.min()) >= 7:
                # one batched decode instead of the reference's per-clip loop: the decoder is causal (left padding, LSTM), so the
                # first lens_i * hop samples of a padded sequence equal the decode of its valid prefix -- as long as the clip is
                # longer than the first conv's reflect padding (6 frames), which looks at the samples after the left edge
                wav = self.vocos.decode_batch(out.transpose(1, 2))
                hop = wav.shape[-1] // n
                out = [wav[i, 0, :int(duration[i]) * hop] for i in range(batch)]
            else:
                audio = []
                for mel, one_mask in zip(out, mask):
                    one = mel[one_mask].transpose(0, 1)[None]
                    audio.append(self.vocos.decode(one).reshape(-1))
                out = audio
        if exists(save_to_filename):                                                 # X3:2289-2303
            import torchaudio
            assert exists(vocoder) or exists(self.vocos)
            assert exists(self.sampling_rate)
            path = Path(save_to_filename)
            path.parents[0].mkdir(exist_ok=True, parents=True)
            for ind, one_audio in enumerate(out):
                name = path.name if len(out) == 1 else f'{ind + 1}.{path.name}'
                torchaudio.save(str(path.parents[0] / name), one_audio.reshape(1, -1).detach().cpu(), sample_rate=self.sampling_rate)
        return out

    @torch.no_grad()
    def velocity(self, x, t, *, clip, context, context_mask=None, roll=None, lens=None, passes=('null',)):
        """All guidance-pass velocities at time t: Float[P, b, n, d] (one `transformer_with_pred_head` per pass, X3:1993)."""
        b, n, _ = x.shape
        if x.shape[-1] != self.num_channels or tuple(clip.shape) != (b, n, self.dim_text) or context.ndim != 3 or context.shape[0] != b \
                or context.shape[-1] != self.dim or (roll is not None and tuple(roll.shape) != (b, n, NOTES)):
            raise ValueError(f'velocity: x {tuple(x.shape)}, clip {tuple(clip.shape)}, context {tuple(context.shape)} do not fit this model '
                             f'(num_channels {self.num_channels}, dim_text {self.dim_text}, dim {self.dim})')
        flags = [0] + [GUIDANCE_PASSES[k] for k in passes]
        order = sorted(range(len(flags)), key=lambda i: (flags[i] & _lib.DROP_CTX) != 0)
        eng = self.engine()
        nc = context.shape[1]
        eng.prepare(b, n, nc, len(flags))
        f32 = lambda v: None if v is None else v.to(device=x.device, dtype=torch.float32).contiguous()
        ctx_lens = _prefix_lens(context_mask, nc, 'context_mask') if context_mask is not None else [nc] * b
        eng.set_conditions(f32(clip), f32(roll), f32(context), [n] * b if lens is None else [int(v) for v in lens], ctx_lens,
                           [flags[i] for i in order])
        pred = eng.forward(f32(x), float(t), len(flags))
        inv = [order.index(i) for i in range(len(flags))]
        return pred[inv]

    def forward(self, *a, **k):
        raise NotImplementedError('the CFM training loss (X3:2307-2588) is outside the sampling path this library implements')
