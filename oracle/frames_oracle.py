"""ORACLE (test infrastructure, not the product): CPU restatement of the reference's piano-roll front end
(src/e2_tts_pytorch/e2_tts_crossatt3.py, "X3"): the static `E2TTS.encode_video_frames` with a frame cache present
(X3:1829-1991) and `E2TTS.encode_frames` around the Video2RollNet call (X3:1525-1555).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

What the reference does (X3 line numbers):
  * 1872-1881  None paths are skipped (they do NOT produce a row); a path may be a tuple (path, start_sample, max_sample)
  * 1885-1907  piano clips read `<video>.generated_frames_raw.2.npz`: arr_0 = grey frames [F, 100, 900, 1] fp32, arr_1 = duration;
               non-piano clips are skipped like None
  * 1928-1940  max_sample defaults to int(duration * 24000); one cached frame per 3 latent frames (960 samples):
                   for i in range(start_sample, max_sample + 960, 960): j = min(round(i / 24000 / (duration / F)), F - 1)
               (Python floats, round-half-even; note F, not F - 1), at most floor(l / 3) + 1 frames
  * 1961-1976  nothing picked -> (None, None); otherwise zero-pad every clip to max(floor(l/3)+1, longest) frames,
               stack and permute to [b', 1, T, 100, 900]
  * 1977-1986  midis = zeros [b', l, 51]
  * 1525-1538  encode_frames: for every frame i the 5 neighbours clamp(i-2..i+2, 0, t-1) -> [b*t, 5, w, h]
  * 1540-1554  net -> sigmoid -> every row repeated 3 times -> cut / zero-pad to l rows
Pinned against the reference's own methods in tests/test_frames_cpu.py (live, where /root/reference exists) and through the
committed fixture tests/golden/frames.npz (oracle/make_golden_frames.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

NOTES = 51
FRAME_SAMPLES = 960          # int(3.0 * 320), X3:1931-1935
SR = 24000


def roll_frame_indices(n_frames: int, duration: float, l: int, start_sample: int = 0, max_sample: int | None = None) -> list[int]:
    """X3:1928-1940: which cached video frame every roll frame copies."""
    if max_sample is None:
        max_sample = int(duration * SR)
    want = math.floor(l / 3.0) + 1
    out = []
    for i in range(start_sample, max_sample + FRAME_SAMPLES, FRAME_SAMPLES):
        out.append(min(round(i / SR / (duration / (n_frames - 0))), n_frames - 1))
        if len(out) >= want:
            break
    return out


def encode_video_frames_cached(clips, l: int, piano: bool = True):
    """clips: list of None or (frames [F, 100, 900, 1] fp32, duration, start_sample, max_sample-or-None).
    Returns (video_frames [b', 1, T, 100, 900], midis [b', l, 51]) as numpy arrays, or (None, None)."""
    picked = []
    for clip in clips:
        if clip is None or not piano:
            continue
        frames, duration, start, max_sample = clip
        idx = roll_frame_indices(frames.shape[0], float(duration), l, start, max_sample)
        picked.append(frames[idx])
    if not picked:
        return None, None
    T = max(math.floor(l / 3.0) + 1, max(p.shape[0] for p in picked))
    out = np.zeros((len(picked), T) + picked[0].shape[1:], dtype=np.float32)
    for b, p in enumerate(picked):
        out[b, :p.shape[0]] = p
    return np.ascontiguousarray(out.transpose(0, 4, 1, 2, 3)), np.zeros((len(picked), l, NOTES), dtype=np.float32)


def frame_windows(x: torch.Tensor) -> torch.Tensor:
    """X3:1530-1538 as written (Python loops): [b, 1, t, w, h] -> [b*t, 5, w, h]."""
    b, c, t, w, h = x.shape
    assert c == 1
    x_all = []
    for i in range(t):
        frames = []
        for j in [-2, -1, 0, 1, 2]:
            f = min(max(i + j, 0), t - 1)
            frames.append(x[:, :, f:f + 1, :, :])
        x_all.append(torch.cat(frames, dim=2))
    return torch.cat(x_all, dim=1).reshape(b * t, 5, w, h)


def encode_frames(x: torch.Tensor, l: int, net) -> torch.Tensor:
    """X3:1525-1555 with `net` standing for self.video2roll_net."""
    b, c, t, w, h = x.shape
    y = torch.sigmoid(net(frame_windows(x)))
    y = y.reshape(b, t, 1, NOTES).repeat(1, 1, 3, 1).reshape(b, t * 3, NOTES)
    d = y.shape[1]
    if d > l:
        y = y[:, :l, :]
    elif d < l:
        y = torch.cat((y, torch.zeros(b, l - d, NOTES)), 1)
    return y
