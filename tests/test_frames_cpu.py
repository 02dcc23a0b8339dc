"""Piano-roll front end, host side: the oracle restatement (oracle/frames_oracle.py) and the drop-in's static
E2TTS.encode_video_frames against the fixture the reference's own methods produced (tests/golden/frames.npz,
oracle/make_golden_frames.py) and, where /root/reference exists, against the reference live."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import frames_oracle as fo, ref_loader, synth
from oracle.make_golden_frames import CASES, ENCODE_FRAMES_CASES

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'frames.npz'))


def _write_cache(tmp, k, seed, F, duration):
    vp = os.path.join(tmp, f'clip{k}.mp4')
    np.savez(vp.replace('.mp4', '.generated_frames_raw.2.npz'), synth.grey_frames(seed, F).numpy(), duration)
    return vp


@pytest.mark.parametrize('k', range(len(CASES)))
def test_oracle_and_dropin_match_reference_fixture(k):
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS
    seed, F, duration, l, start, max_sample = CASES[k]
    assert list(GOLD[f'meta{k}']) == [seed, F, duration, l, start, -1 if max_sample is None else max_sample]
    idx = fo.roll_frame_indices(F, duration, l, start, max_sample)
    gold_idx = GOLD[f'idx{k}'].tolist()
    want = l // 3 + 1
    assert len(gold_idx) == max(want, len(idx))
    assert gold_idx[:len(idx)] == idx and all(v == -1 for v in gold_idx[len(idx):])
    frames = synth.grey_frames(seed, F).numpy()
    vf, midis = fo.encode_video_frames_cached([(frames, duration, start, max_sample)], l)
    assert np.array_equal(vf[0, 0].astype(np.float64).sum((1, 2)), GOLD[f'sum{k}'])
    with tempfile.TemporaryDirectory() as tmp:
        vp = _write_cache(tmp, k, seed, F, duration)
        ours_vf, ours_midis = E2TTS.encode_video_frames([vp if (start == 0 and max_sample is None) else (vp, start, max_sample)], l, True)
    assert ours_vf.dtype == torch.float32 and ours_vf.device.type == 'cpu'            # the reference returns host tensors
    assert np.array_equal(ours_vf.numpy(), vf) and np.array_equal(ours_midis.numpy(), midis)


def test_dropin_batch_semantics_match_reference_fixture():
    """None clips are dropped (not zero rows), shorter clips are zero-padded, no piano clip -> (None, None); 2-argument call
    (the dataset collate's form, trainer3:1377) works."""
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS
    with tempfile.TemporaryDirectory() as tmp:
        p0 = _write_cache(tmp, 0, *CASES[0][:3])
        p1 = _write_cache(tmp, 1, *CASES[1][:3])
        vf, midis = E2TTS.encode_video_frames([p1, None, p0], 225, True)
        assert list(vf.shape) == GOLD['batch_shape'].tolist() and tuple(midis.shape) == (2, 225, 51)
        assert np.array_equal(vf.numpy().astype(np.float64).sum((1, 3, 4)), GOLD['batch_sum'])
        assert E2TTS.encode_video_frames([None, None], 100, True) == (None, None)
        assert E2TTS.encode_video_frames([p0], 100, False) == (None, None)
        vf2, _ = E2TTS.encode_video_frames([p0], 225)
        assert torch.equal(vf2[0], vf[1])
        with pytest.raises(RuntimeError, match='no cached frames'):
            E2TTS.encode_video_frames([os.path.join(tmp, 'missing.mp4')], 100, True)


@pytest.mark.parametrize('k', range(len(ENCODE_FRAMES_CASES)))
def test_encode_frames_oracle_matches_reference_fixture(k):
    seed, b, t, l = ENCODE_FRAMES_CASES[k]
    x = torch.stack([synth.grey_frames(seed + 100 * i, t)[..., 0] for i in range(b)])[:, None]
    roll = fo.encode_frames(x, l, synth.StandInRollNet())
    assert np.array_equal(roll.numpy(), GOLD[f'roll{k}'])
    assert roll.std() > 0.1                                   # the stand-in net is live


@pytest.mark.skipif(not ref_loader.reference_available(), reason='/root/reference not present')
def test_live_reference_agrees():
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS
    x3 = ref_loader.load_x3()
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for k, (seed, F, duration, l, start, max_sample) in enumerate(CASES):
            vp = _write_cache(tmp, k, seed, F, duration)
            paths.append(vp if (start == 0 and max_sample is None) else (vp, start, max_sample))
        for l in (90, 225, 301):
            ref_vf, ref_m = x3.E2TTS.encode_video_frames(paths + [None], l, True)
            our_vf, our_m = E2TTS.encode_video_frames(paths + [None], l, True)
            assert torch.equal(ref_vf, our_vf) and torch.equal(ref_m, our_m)
    x = synth.grey_frames(42, 7)[..., 0][None, None]
    assert torch.equal(fo.frame_windows(x)[3, 0], x[0, 0, 1]) and torch.equal(fo.frame_windows(x)[0, 0], x[0, 0, 0])
