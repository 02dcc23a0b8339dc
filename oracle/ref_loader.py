"""TEST INFRASTRUCTURE ONLY -- import the reference's own X3 module through shims.

Imports /root/reference/src/e2_tts_pytorch/e2_tts_crossatt3.py *verbatim*
(nothing is copied into this repo) by pre-populating sys.modules with

  * functional restatements of the missing pinned dependencies
    (oracle/third_party.py: x_transformers, torchdiffeq, einx), and
  * inert stubs for modules the sampling path never calls
    (vocos, g2p_en, jieba, pypinyin, moviepy.editor, audioldm.*).

Everything in X3 (sampler, CFG, transformer wiring, masks, conv, AdaLN,
cross-condition) is then the reference's own code.  /root/reference exists only
in the build container, never on the GPU box: this loader is used to pin
oracle/e2_oracle.py (the restatement that travels) and to generate the golden
vectors under tests/golden/ (oracle/make_golden.py).
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.machinery
import io
import os
import sys
import types

from . import third_party as tp

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')


def _pick_root():
    """The reference tree itself (build container) or the files oracle/stage_reference.py staged under the git-ignored
    baseline/_ref/ (what travels to the GPU box)."""
    env = os.environ.get('E2B_REFERENCE_ROOT')
    for root in ([env] if env else []) + ['/root/reference', _STAGED]:
        if os.path.isfile(os.path.join(root, 'src', 'e2_tts_pytorch', 'e2_tts_crossatt3.py')):
            return root
    return env or '/root/reference'


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'src', 'e2_tts_pytorch', 'e2_tts_crossatt3.py'))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Inert:
    def __init__(self, *a, **k):
        raise RuntimeError('inert oracle stub: this dependency is outside the CFM sampling path')


_X3 = None


def load_x3():
    """Return the reference module e2_tts_pytorch.e2_tts_crossatt3 (imported once)."""
    global _X3
    if _X3 is not None:
        return _X3
    if not reference_available():
        raise FileNotFoundError(f'reference not found under {REFERENCE_ROOT}')

    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k.split('.')[0] in ('e2_tts_pytorch', 'einx', 'torchdiffeq', 'x_transformers', 'vocos', 'g2p_en',
                                    'jieba', 'pypinyin', 'moviepy', 'audioldm', 'Video2RollNet')}
    for k in saved:
        del sys.modules[k]

    # functional shims
    _mod('einx', less=tp.einx.less, greater_equal=tp.einx.greater_equal, where=tp.einx.where,
         multiply=tp.einx.multiply, divide=tp.einx.divide)
    _mod('torchdiffeq', odeint=tp.odeint)
    xt = _mod('x_transformers', Attention=tp.Attention, FeedForward=tp.FeedForward, RMSNorm=tp.RMSNorm,
              AdaptiveRMSNorm=tp.AdaptiveRMSNorm)
    xt.x_transformers = _mod('x_transformers.x_transformers', RotaryEmbedding=tp.RotaryEmbedding)

    # inert stubs
    _mod('vocos', Vocos=_Inert)
    _mod('g2p_en', G2p=_Inert)
    _mod('jieba', cut=_Inert)
    _mod('pypinyin', lazy_pinyin=_Inert, Style=_Inert)
    mp = _mod('moviepy')
    mp.editor = _mod('moviepy.editor', AudioFileClip=_Inert, VideoFileClip=_Inert)
    al = _mod('audioldm')
    al.audio = _mod('audioldm.audio')
    al.audio.stft = _mod('audioldm.audio.stft', TacotronSTFT=_Inert)
    al.variational_autoencoder = _mod('audioldm.variational_autoencoder', AutoencoderKL=_Inert)
    al.utils = _mod('audioldm.utils', default_audioldm_config=_Inert, get_metadata=_Inert)

    src = os.path.join(REFERENCE_ROOT, 'src')
    sys.path[:0] = [src, os.path.join(src, 'audeo')]
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            _X3 = importlib.import_module('e2_tts_pytorch.e2_tts_crossatt3')
    finally:
        sys.path.remove(src)
        sys.path.remove(os.path.join(src, 'audeo'))
        # leave the e2_tts_pytorch name free for the repo's own drop-in package
        for k in [k for k in sys.modules if k.split('.')[0] == 'e2_tts_pytorch']:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None and k.split('.')[0] == 'e2_tts_pytorch':
                sys.modules[k] = v
    return _X3


SHIPPED_TRANSFORMER = dict(depth=12, dim=1024, dim_text=1280, heads=16, dim_head=64, if_text_modules=True,
                           if_cross_attn=True, if_audio_conv=True, if_text_conv=True)


def build_reference_model(transformer: dict | None = None, seed: int = 0, num_channels: int = 128,
                          if_cond_proj_in: bool = False):
    """E2TTS as src/inference_v2a.py:74-110 builds it, minus the pretrained encoders (no weights offline)."""
    import torch
    x3 = load_x3()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = x3.E2TTS(
            duration_predictor=None,
            transformer=dict(transformer or SHIPPED_TRANSFORMER),
            tokenizer='char_utf8',
            audiocond_drop_prob=1.1, cond_drop_prob=-0.1, prompt_drop_prob=-0.1,
            if_cond_proj_in=if_cond_proj_in, if_embed_text=False, if_text_encoder2=False, if_clip_encoder=False,
            num_channels=num_channels, sampling_rate=24000,
        )
    m.vocos = None
    m.eval()
    return m


def reference_sample(m, *, y0, clip, ctx, ctx_mask, frames_embed=None, lens=None, steps=32, cfg_strength=2.0,
                     remove_parallel_component=False, sway_sampling=True, cond=None, cond_lens=None, audio_drop_prompt=None):
    """Run the reference's own E2TTS.sample on injected conditions.

    y0 replaces the torch.randn_like draw at X3:2248 (the only RNG use inside sample()); the CLIP stream enters as
    the float `text=` tensor (X3:2040); T5 output enters by overriding encode_text (X3:2057); a precomputed piano-roll
    enters by overriding encode_frames (X3:2170).  With `cond` / `cond_lens` the call is the in-painting mode: the
    reference's `lens` = cond_lens < `duration` = lens (X3:2196-2228, 2259-2260).
    """
    import torch
    b, n, _ = y0.shape
    if lens is None:
        lens = torch.full((b,), n, dtype=torch.long)
    m.encode_text = lambda prompt: (ctx.clone(), ctx_mask.clone())
    frames_arg = None
    if frames_embed is not None:
        m.encode_frames = lambda frames, l: frames_embed.clone()
        frames_arg = torch.zeros(b, 1, 1, 1, 1)
    real_randn_like = torch.randn_like
    calls = []

    def fake_randn_like(t, *a, **k):
        calls.append(1)
        assert tuple(t.shape) == tuple(y0.shape)
        return y0.clone().to(t.dtype)

    torch.randn_like = fake_randn_like
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            out = m.sample(
                cond=torch.zeros(b, n, y0.shape[-1]) if cond is None else cond.clone(), text=clip.clone(), duration=lens.clone(),
                lens=lens.clone() if cond_lens is None else cond_lens.clone(),
                steps=steps, cfg_strength=cfg_strength, remove_parallel_component=remove_parallel_component,
                sway_sampling=sway_sampling, prompt=['the sound of'] * b, video_drop_prompt=[False] * b,
                audio_drop_prompt=audio_drop_prompt, frames=frames_arg, return_raw_output=True)
    finally:
        torch.randn_like = real_randn_like
    assert len(calls) == 1
    return out
