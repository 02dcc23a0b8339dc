"""Generates tests/golden/frames.npz: inputs and outputs of the REFERENCE's own piano-roll front end -- the static
E2TTS.encode_video_frames (X3:1829-1991) run on frame caches written to a temporary directory, and E2TTS.encode_frames
(X3:1525-1555) run with oracle.synth.StandInRollNet in place of the pretrained Video2RollNet.  Needs /root/reference (this
container only); the fixture (indices, shapes, checksums, the small roll tensors) travels."""
import os, sys, tempfile
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, synth

# (clip seed, cached frames F, duration s, l latent frames, start_sample, max_sample)
CASES = [(0, 25, 3.0, 225, 0, None), (1, 11, 1.37, 120, 0, None), (2, 40, 2.0, 90, 0, None), (3, 30, 4.0, 200, 9600, 60000), (4, 2, 0.3, 30, 0, None)]
ENCODE_FRAMES_CASES = [(10, 2, 9, 30), (11, 1, 5, 11), (12, 1, 6, 18)]     # (seed, b, t, l): l > 3t pads, l < 3t cuts, l == 3t


def main():
    x3 = ref_loader.load_x3()
    store = {}
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for k, (seed, F, duration, l, start, max_sample) in enumerate(CASES):
            frames = synth.grey_frames(seed, F).numpy()
            vp = os.path.join(tmp, f'clip{k}.mp4')
            np.savez(vp.replace('.mp4', '.generated_frames_raw.2.npz'), frames, duration)      # the format X3:1905 writes
            arg = vp if (start == 0 and max_sample is None) else (vp, start, max_sample)
            paths.append(arg)
            vf, midis = x3.E2TTS.encode_video_frames([arg], l, True)
            vf = vf.numpy()
            assert vf.shape[:2] == (1, 1) and vf.shape[3:] == (100, 900) and tuple(midis.shape) == (1, l, 51) and not midis.any()
            idx = []
            for t in range(vf.shape[2]):                       # frames are copies: the probe pixel recovers which one was picked
                hit = np.nonzero((frames[:, :, :, 0] == vf[0, 0, t][None]).all((1, 2)))[0]
                idx.append(int(hit[0]) if len(hit) else -1)    # -1 = zero padding
            store[f'idx{k}'] = np.asarray(idx, dtype=np.int32)
            store[f'meta{k}'] = np.asarray([seed, F, duration, l, start, -1 if max_sample is None else max_sample], dtype=np.float64)
            store[f'sum{k}'] = vf[0, 0].astype(np.float64).sum((1, 2))
        # a batch: None and a second clip; None rows are dropped, the shorter clip is zero-padded (X3:1872-1876, 1966-1975)
        vf, midis = x3.E2TTS.encode_video_frames([paths[1], None, paths[0]], 225, True)
        store['batch_shape'] = np.asarray(vf.shape)
        store['batch_sum'] = vf.numpy().astype(np.float64).sum((1, 3, 4))
        assert x3.E2TTS.encode_video_frames([None, None], 100, True) == (None, None)
        assert x3.E2TTS.encode_video_frames([paths[0]], 100, False) == (None, None)
    m = ref_loader.build_reference_model(transformer=dict(ref_loader.SHIPPED_TRANSFORMER, depth=2))
    m.video2roll_net = synth.StandInRollNet()
    for k, (seed, b, t, l) in enumerate(ENCODE_FRAMES_CASES):
        x = torch.stack([synth.grey_frames(seed + 100 * i, t)[..., 0] for i in range(b)])[:, None]      # [b, 1, t, 100, 900]
        with torch.no_grad():
            roll = m.encode_frames(x, l)
        assert tuple(roll.shape) == (b, l, 51)
        store[f'roll{k}'] = roll.numpy()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'frames.npz'), **store)
    print('wrote', len(CASES), '+', len(ENCODE_FRAMES_CASES), 'cases')


if __name__ == '__main__':
    main()
