"""Generates tests/golden/staging.npz: inputs and outputs of the REFERENCE's own E2TTS.encode_video (X3:1659-1827) run on
feature caches written to a temporary directory.  Needs /root/reference (this container only); the fixture travels."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader


def cases():
    rng = np.random.default_rng(7)
    d = 1280
    out = []
    for F, duration, l, start, max_sample in [(300, 10.0, 750, 0, None), (37, 3.21, 750, 0, None), (251, 8.37, 400, 0, None),
                                               (120, 29.97, 2250, 0, None), (90, 6.0, 300, 4800, 100000), (2, 0.5, 50, 0, None)]:
        out.append((rng.standard_normal((F, d)).astype(np.float32), duration, l, start, max_sample))
    return out


def main():
    m = ref_loader.build_reference_model(transformer=dict(ref_loader.SHIPPED_TRANSFORMER, depth=2))
    store = {}
    with tempfile.TemporaryDirectory() as tmp:
        for k, (emb, duration, l, start, max_sample) in enumerate(cases()):
            vp = os.path.join(tmp, f'clip{k}.mp4')
            np.savez(vp.replace('.mp4', '.generated.npz'), emb, duration)          # the format X3:1796 writes
            arg = vp if (start == 0 and max_sample is None) else (vp, start, max_sample)
            ref = m.encode_video([arg, None], l).cpu().numpy()
            assert ref.shape == (2, l, 1280) and not ref[1].any()
            # the rows are copies, so the row index recovers exactly which frame the reference picked
            idx = []
            for r in ref[0]:
                hit = np.nonzero((emb == r[None, :]).all(1))[0]
                idx.append(int(hit[0]) if len(hit) else -1)
            store[f'idx{k}'] = np.asarray(idx, dtype=np.int32)
            store[f'meta{k}'] = np.asarray([emb.shape[0], duration, l, start, -1 if max_sample is None else max_sample], dtype=np.float64)
            store[f'sum{k}'] = ref[0].astype(np.float64).sum(1)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'staging.npz'), **store)
    print('wrote', len(cases()), 'cases')


if __name__ == '__main__':
    main()
