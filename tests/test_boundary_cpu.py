"""Host-side checks that need no GPU: the C-ABI library builds/loads and exports every symbol include/e2b.h declares, the
drop-in classes keep the reference's names / kwargs / state-dict keys, and the product path refuses to run without CUDA
(there is no CPU fallback)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def libpath():
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'))
    import build
    return build.build()


def test_library_exports_every_declared_symbol(libpath):
    hdr = open(os.path.join(ROOT, 'include', 'e2b.h')).read()
    declared = set(re.findall(r'\b(e2b_[a-z_0-9]+)\s*\(', hdr))
    assert {'e2b_create', 'e2b_load_weights', 'e2b_prepare', 'e2b_set_conditions', 'e2b_forward', 'e2b_sample',
            'e2b_transformer_forward', 'e2b_guided_euler', 'e2b_melspec', 'e2b_last_error', 'e2b_destroy'} <= declared
    lib = ctypes.CDLL(libpath)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/e2b.h but not exported'
    from e2_tts_pytorch import _lib
    for name in _lib.EXPORTED_SYMBOLS:
        assert hasattr(lib, name), name


def test_library_is_sm100a_tcgen05_tma(libpath):
    """SASS evidence: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA) -- and no legacy HMMA path."""
    import shutil, subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', libpath], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    assert re.search(r'UTC\w*MMA', sass) and 'LDTM' in sass and 'UTMALDG' in sass
    assert not re.search(r'\bHMMA\b', sass)


def test_state_dict_keys_match_reference_names():
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS
    cfg = synth.TINY
    m = E2TTS(transformer=dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'],
                               heads=cfg['heads'], dim_head=64, max_seq_len=cfg['max_seq_len'], if_text_conv=True),
              duration_predictor=None, if_cond_proj_in=False, if_embed_text=False, if_text_encoder2=False, if_clip_encoder=False,
              num_channels=cfg['num_channels'], sampling_rate=24000, audiocond_drop_prob=1.1, cond_drop_prob=-0.1,
              prompt_drop_prob=-0.1, tokenizer='phoneme_zh')
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in synth.random_state_dict(**cfg).items()}     # pinned against X3 in test_oracle_vs_reference
    assert ours == ref
    assert m.num_channels == cfg['num_channels'] and m.sampling_rate == 24000 and m.vocos is None
    # default init keeps the reference's dead-at-init groups
    sd = m.state_dict()
    assert torch.count_nonzero(sd['transformer.layers.0.0.2.to_gamma.weight']) == 0
    assert torch.all(sd['transformer.layers.0.0.4.to_gamma.bias'] == -2)
    assert torch.all(sd['transformer.layers.0.0.3.to_v_head_gate.bias'] == 10)


def test_sample_signature_keeps_reference_kwargs():
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS, Transformer, MelSpec, EncodecWrapper, DurationPredictor  # noqa: F401
    ref_kwargs = ['text', 'lens', 'duration', 'steps', 'cfg_strength', 'remove_parallel_component', 'sway_sampling', 'max_duration',
                  'vocoder', 'return_raw_output', 'save_to_filename', 'prompt', 'video_drop_prompt', 'audio_drop_prompt', 'video_paths',
                  'frames', 'midis']                                                    # e2_tts_crossatt3.py:2128-2148
    sig = inspect.signature(E2TTS.sample)
    for k in ref_kwargs:
        assert k in sig.parameters, k
    assert sig.parameters['steps'].default == 32 and sig.parameters['cfg_strength'].default == 1.
    assert sig.parameters['remove_parallel_component'].default is True and sig.parameters['max_duration'].default == 4096
    fsig = inspect.signature(Transformer.forward)
    assert list(fsig.parameters)[1:] == ['x', 'times', 'mask', 'text_embed', 'frames_embed', 'context', 'context_mask']


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS, MelSpec
    cfg = synth.TINY
    m = E2TTS(transformer=dict(depth=2, dim=128, dim_text=128, dim_frames=64, heads=2, dim_head=64, max_seq_len=64, if_text_conv=True),
              if_cond_proj_in=False, if_embed_text=False, if_text_encoder2=False, num_channels=64)
    with pytest.raises(RuntimeError, match='CUDA'):
        m.sample(torch.zeros(1, 8, 64), text=torch.zeros(1, 8, 128), context=torch.zeros(1, 4, 128), return_raw_output=True)
    with pytest.raises(RuntimeError, match='CUDA'):
        MelSpec()(torch.zeros(1, 4096))


def test_mel_filterbank_matches_torchaudio():
    import torchaudio
    from e2_tts_pytorch.e2_tts_crossatt3 import MelSpec
    ref = torchaudio.functional.melscale_fbanks(513, 0., 12000., 100, 24000, norm=None, mel_scale='htk')
    assert torch.allclose(MelSpec.mel_filterbank(513, 0., 12000., 100, 24000), ref, atol=1e-6)


def test_prompt_context_is_encoded_once_per_distinct_prompt():
    """encode_text_cached: distinct prompts are encoded once, repeated calls hit the cache, and the assembled
    batch equals the jointly encoded, right-padded batch on every valid token."""
    import torch
    from oracle import synth
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS
    cfg = synth.TINY
    tr = dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'], heads=cfg['heads'], dim_head=64,
              max_seq_len=cfg['max_seq_len'], if_text_modules=True, if_cross_attn=True, if_audio_conv=True, if_text_conv=True)
    m = E2TTS(duration_predictor=None, transformer=tr, tokenizer='char_utf8', audiocond_drop_prob=1.1, cond_drop_prob=-0.1,
              prompt_drop_prob=-0.1, if_cond_proj_in=False, if_embed_text=False, if_text_encoder2=False, if_clip_encoder=False,
              num_channels=cfg['num_channels'], sampling_rate=24000)
    calls = []

    def fake_encode_text(prompts):                       # stands in for the frozen T5: one row per word, value = word hash
        calls.append(list(prompts))
        toks = [[float(sum(map(ord, w))) for w in p.split()] for p in prompts]
        nc = max(map(len, toks))
        hidden = torch.zeros(len(prompts), nc, cfg['dim'])
        mask = torch.zeros(len(prompts), nc, dtype=torch.bool)
        for i, t in enumerate(toks):
            hidden[i, :len(t)] = torch.tensor(t)[:, None] + torch.arange(cfg['dim'])[None, :]
            hidden[i, len(t):] = -7.0                    # garbage in the padded positions, as a real encoder leaves it
            mask[i, :len(t)] = True
        return hidden, mask

    m.encode_text = fake_encode_text
    batch = ['the sound of rain', 'the sound of playing piano', 'the sound of rain', 'a dog']
    ctx, msk = m.encode_text_cached(batch)
    assert calls == [['the sound of rain', 'the sound of playing piano', 'a dog']]
    joint, jmask = fake_encode_text(batch)
    assert torch.equal(msk, jmask) and torch.equal(ctx[msk], joint[jmask]) and not ctx[~msk].any()
    ctx2, msk2 = m.encode_text_cached(['a dog', 'the sound of rain'])
    assert len(calls) == 2                               # only the joint reference call above was added: both prompts were cached
    assert ctx2.shape == (2, 4, cfg['dim']) and msk2.sum().item() == 6
    m.encode_text_cached(['something new'])
    assert calls[-1] == ['something new']


def test_config_struct_matches_header_and_integration_snippet(libpath):
    """e2b_config as declared in include/e2b.h, as bound in _lib.Config, as the built library reports it (e2b_config_size) and
    as INTEGRATION.md's Level-2 snippet declares it must all agree -- and e2b_create must get through every field check with the
    snippet's struct (on a box without a GPU it then stops at "no CUDA device"; with one it succeeds)."""
    from e2_tts_pytorch import _lib
    hdr = open(os.path.join(ROOT, 'include', 'e2b.h')).read()
    body = re.search(r'typedef struct e2b_config \{(.*?)\} e2b_config;', hdr, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = [f.strip() for decl in re.findall(r'int ([^;]+);', body) for f in decl.split(',')]
    assert fields == [f for f, _ in _lib.Config._fields_]
    lib = ctypes.CDLL(libpath)
    assert lib.e2b_config_size() == ctypes.sizeof(_lib.Config) == 4 * len(fields)

    md = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    snippet = re.search(r'(class _Cfg\(C\.Structure\):.*?\n)class _T', md, re.S).group(1)
    ns = {'C': ctypes, '_e2b': lib}
    exec(snippet, ns)                                        # also runs the snippet's own e2b_config_size assertion
    assert [f for f, _ in ns['_Cfg']._fields_] == fields
    call = re.search(r'cfg = (_Cfg\(.*?\))\s*#', md, re.S).group(1)

    class _Tr:
        depth, dim, dim_text, dim_frames, num_registers, max_seq_len = 12, 1024, 1280, 512, 32, 8192

    class _Self:
        num_channels = 128
    cfg = eval(call, dict(ns, t=_Tr, self=_Self))
    assert cfg.precision == 0 and cfg.ff_mult == 4
    lib.e2b_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.e2b_last_error.restype = ctypes.c_char_p
    lib.e2b_last_error.argtypes = [ctypes.c_void_p]
    h = ctypes.c_void_p()
    rc = lib.e2b_create(ctypes.byref(cfg), ctypes.byref(h))
    if torch.cuda.is_available():
        assert rc == 0 and h
        lib.e2b_destroy.argtypes = [ctypes.c_void_p]
        lib.e2b_destroy(h)
    else:
        assert rc != 0 and b'no CUDA device' in lib.e2b_last_error(None)
    cfg.precision = 7                                        # a garbage 14th field (what a 13-int struct would hand over) is refused
    assert lib.e2b_create(ctypes.byref(cfg), ctypes.byref(h)) != 0 and b'precision' in lib.e2b_last_error(None)


def test_app_import_line_and_static_frame_call_resolve():
    """app.py:48-49 / predict.py:48-49 import line and the static call at app.py:236 resolve against the drop-in."""
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS, DurationPredictor, MelSpec, EncodecWrapper  # noqa: F401
    assert isinstance(inspect.getattr_static(E2TTS, 'encode_video_frames'), staticmethod)
    assert list(inspect.signature(E2TTS.encode_video_frames).parameters) == ['video_paths', 'l', 'piano']
    assert E2TTS.encode_video_frames([None], 75, True) == (None, None)                # X3:1961-1962
    for name in ('encode_video', 'encode_text', 'encode_frames', 'sample', 'transformer_with_pred_head'):
        assert name in ('transformer_with_pred_head',) or callable(getattr(E2TTS, name))


def test_feature_cache_paths_follow_the_reference():
    """X3:1679-1704: videos under the dataset root map /video/ -> /feature*/ + .npz, everything else -> .generated*.npz."""
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS

    class _M:
        video_encoder = 'clip_vit'
        VIDEO_FEATURE_SUFFIX, VGGSOUND_ROOT, VGGSOUND_FEATURE_DIR = E2TTS.VIDEO_FEATURE_SUFFIX, E2TTS.VGGSOUND_ROOT, E2TTS.VGGSOUND_FEATURE_DIR
    f = E2TTS.video_feature_path
    assert f(_M, '/data/clips/a.mp4') == '/data/clips/a.generated.npz'
    assert f(_M, '/ailab-train2/speech/zhanghaomin/VGGSound/video/x_000030.mp4') == '/ailab-train2/speech/zhanghaomin/VGGSound/feature/x_000030.npz'
    _M.video_encoder = 'dinov2'
    assert f(_M, '/ailab-train2/speech/zhanghaomin/VGGSound/video/x.mp4') == '/ailab-train2/speech/zhanghaomin/VGGSound/feature_dinov2/x.npz'
    assert f(_M, 'b.mp4') == 'b.generated.dinov2.npz'
    _M.video_encoder = 'nope'
    with pytest.raises(Exception, match='Invalid video_encoder'):
        f(_M, 'b.mp4')
