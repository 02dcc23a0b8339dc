"""Condition staging on the GPU (csrc/staging.cu via E2TTS.encode_video) against the oracle: bit-exact, it is index + copy work."""
import numpy as np
import pytest
import torch

from gpu_util import DEV
from oracle import staging_oracle as so, synth
from oracle.make_golden_staging import cases
from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS

pytestmark = pytest.mark.gpu


def _model():
    cfg = synth.TINY
    tr = dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=1280, dim_frames=cfg['dim_frames'], heads=cfg['heads'], dim_head=64,
              max_seq_len=cfg['max_seq_len'], if_text_modules=True, if_cross_attn=True, if_audio_conv=True, if_text_conv=True)
    return E2TTS(duration_predictor=None, transformer=tr, tokenizer='char_utf8', audiocond_drop_prob=1.1, cond_drop_prob=-0.1,
                 prompt_drop_prob=-0.1, if_cond_proj_in=False, if_embed_text=False, if_text_encoder2=False, if_clip_encoder=False,
                 num_channels=cfg['num_channels'], sampling_rate=24000).to(DEV)


def _write(tmp_path, k, emb, duration, suffix='.generated.npz'):
    vp = str(tmp_path / f'clip{k}.mp4')
    np.savez(vp.replace('.mp4', suffix), emb, duration)
    return vp


def test_encode_video_matches_oracle_on_reference_fixture_cases(tmp_path):
    m = _model()
    g = np.load(__file__.replace('test_gpu_7_staging.py', 'golden/staging.npz'))
    for k, (emb, duration, l, start, max_sample) in enumerate(cases()):
        vp = _write(tmp_path, k, emb, duration)
        arg = vp if (start == 0 and max_sample is None) else (vp, start, max_sample)
        out = m.encode_video([arg, None], l)
        assert out.shape == (2, l, 1280) and out.dtype == torch.float32 and out.is_cuda
        ref = so.encode_video_cached([(emb, duration, start, max_sample), None], l, 1280)
        assert np.array_equal(out.cpu().numpy(), ref)
        assert np.array_equal(out[0].double().sum(1).cpu().numpy(), g[f'sum{k}'])      # the reference's own output


def test_encode_video_random_batches_and_half_even(tmp_path):
    m = _model()
    rng = np.random.default_rng(11)
    paths, clips = [], []
    for k in range(24):
        F = int(rng.integers(2, 400))
        duration = float(rng.uniform(0.2, 31.0))
        emb = rng.standard_normal((F, 1280)).astype(np.float32)
        start = int(rng.integers(0, 3)) * 320 * int(rng.integers(0, 50))
        max_sample = None if k % 3 else start + int(rng.integers(1, 400000))
        vp = _write(tmp_path, k, emb, duration)
        paths.append(vp if (start == 0 and max_sample is None) else (vp, start, max_sample))
        clips.append((emb, duration, start, max_sample))
        if k % 7 == 3:
            paths.append(None)
            clips.append(None)
    for l in (1, 375, 750, 2250):
        out = m.encode_video(paths, l)
        assert np.array_equal(out.cpu().numpy(), so.encode_video_cached(clips, l, 1280))
    # exact .5 positions (see tests/test_oracle_staging.py): round half to even on the device too
    m2 = _model()
    m2.sampling_rate, m2.frame_size = 16384, 256
    emb = rng.standard_normal((5, 1280)).astype(np.float32)
    vp = _write(tmp_path, 99, emb, 0.0625)
    out = m2.encode_video([(vp, 0, 10 ** 9)], 4).cpu().numpy()
    assert np.array_equal(out[0], emb[[0, 2, 2, 4]])


def test_encode_video_errors_and_sample_entry(tmp_path):
    m = _model()
    with pytest.raises(RuntimeError, match='no cached video features'):
        m.encode_video([str(tmp_path / 'missing.mp4')], 10)
    vp = _write(tmp_path, 0, np.zeros((1, 1280), np.float32), 1.0)
    with pytest.raises(ZeroDivisionError):
        m.encode_video([vp], 10)
    vp = _write(tmp_path, 1, np.zeros((4, 640), np.float32), 1.0)
    with pytest.raises(RuntimeError, match='expected embeddings'):
        m.encode_video([vp], 10)
    vp = _write(tmp_path, 2, np.ones((4, 1280), np.float32), 1.0)
    with pytest.raises(RuntimeError, match='non-empty'):
        m.encode_video([(vp, 24000, 24000)], 10)
    m.video_encoder = 'bogus'
    with pytest.raises(Exception, match='Invalid video_encoder'):
        m.encode_video([vp], 10)
    m.video_encoder = 'dinov2'
    vp = _write(tmp_path, 3, np.ones((4, 1280), np.float32), 1.0, '.generated.dinov2.npz')
    assert m.encode_video([vp], 10).shape == (1, 10, 1280)
