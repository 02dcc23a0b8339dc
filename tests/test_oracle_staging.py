"""Condition staging: the numpy oracle against the committed fixture made from the REFERENCE's own
E2TTS.encode_video (oracle/make_golden_staging.py), and -- where /root/reference exists -- against that method run live."""
import os

import numpy as np
import pytest

from oracle import ref_loader, staging_oracle as so
from oracle.make_golden_staging import cases

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'staging.npz')


def test_oracle_matches_reference_fixture():
    g = np.load(GOLDEN)
    for k, (emb, duration, l, start, max_sample) in enumerate(cases()):
        meta = g[f'meta{k}']
        assert (int(meta[0]), float(meta[1]), int(meta[2]), int(meta[3])) == (emb.shape[0], duration, l, start)
        idx = so.frame_indices(emb.shape[0], duration, l, 24000, 320, start, max_sample)
        ref = g[f'idx{k}']
        assert ref.shape == (l,)
        assert (ref[:len(idx)] == np.asarray(idx)).all()            # the very frames the reference copied
        assert (ref[len(idx):] == -1).all()                         # then zero padding
        out = so.encode_video_cached([(emb, duration, start, max_sample), None], l, 1280)
        assert not out[1].any()
        assert np.array_equal(out[0].astype(np.float64).sum(1), g[f'sum{k}'])


def test_frame_index_rule_details():
    # round-half-even (Python round, X3:1806): with sr = 2^14, frame 256, duration / (F - 1) = 1/64 every position is exactly
    # k + 0.5 in binary floating point -> 0.5 -> 0, 1.5 -> 2, 2.5 -> 2, 3.5 -> 4
    assert so.frame_indices(5, 0.0625, 4, 16384, 256, 0, 10 ** 9) == [0, 2, 2, 4]
    F, sr, fs = 11, 24000, 320
    # the index saturates at F - 1 and the list stops at l or at max_sample
    assert so.frame_indices(5, 1.0, 1000, sr, fs)[-1] == 4
    assert len(so.frame_indices(5, 1.0, 1000, sr, fs)) == len(range(0, 24000, 320))
    assert len(so.frame_indices(5, 1.0, 10, sr, fs)) == 10
    assert so.frame_indices(5, 1.0, 10, sr, fs, 24000, 24000) == []
    with pytest.raises(ZeroDivisionError):
        so.frame_indices(1, 1.0, 10, sr, fs)
    assert so.feature_path('/a/b.mp4') == '/a/b.generated.npz'
    assert so.feature_path('/a/b.mp4', 'dinov2') == '/a/b.generated.dinov2.npz'
    assert so.split_path(('/a/b.mp4', 5, 9)) == ('/a/b.mp4', 5, 9) and so.split_path('/a/b.mp4') == ('/a/b.mp4', 0, None)


@pytest.mark.skipif(not ref_loader.reference_available(), reason='/root/reference not present')
def test_oracle_matches_reference_live(tmp_path):
    m = ref_loader.build_reference_model(transformer=dict(ref_loader.SHIPPED_TRANSFORMER, depth=2))
    rng = np.random.default_rng(3)
    paths, clips = [], []
    for k, (F, duration, start, max_sample) in enumerate([(77, 4.4, 0, None), (300, 10.0, 3200, 200000), (9, 2.0, 0, None)]):
        emb = rng.standard_normal((F, 1280)).astype(np.float32)
        vp = str(tmp_path / f'v{k}.mp4')
        np.savez(so.feature_path(vp), emb, duration)
        paths.append(vp if (start == 0 and max_sample is None) else (vp, start, max_sample))
        clips.append((emb, duration, start, max_sample))
    paths.insert(1, None)
    clips.insert(1, None)
    ref = m.encode_video(paths, 500).cpu().numpy()
    assert np.array_equal(ref, so.encode_video_cached(clips, 500, 1280))
