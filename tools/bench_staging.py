"""Time the condition-staging gather (csrc/staging.cu) at the C2 shape: 64 clips x 750 latent frames x 1280 channels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch
from e2_tts_pytorch import _lib
DEV = 'cuda:0'
B, l, d, F = 64, 750, 1280, 300
emb = torch.randn(B * F, d, device=DEV)
meta = torch.tensor([(b * F, F, l, 0) for b in range(B)], dtype=torch.int64, device=DEV)
dur = torch.full((B,), 10.0, dtype=torch.float64, device=DEV)
out = torch.empty(B, l, d, device=DEV)
run = lambda: _lib.check(_lib.lib().e2b_stage_clip(_lib.ptr(emb), _lib.ptr(meta), _lib.ptr(dur), B, l, d, 24000, 320, _lib.ptr(out),
                                                  _lib.stream_ptr()), None, 'stage')
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 50
print(f'stage_clip B={B} l={l} d={d}: {us:.1f} us, {8.0 * B * l * d / us / 1e3:.0f} GB/s algorithmic (4 B read + 4 B written per element; '
      f'{4.0 * B * (l + F) * d / us / 1e3:.0f} GB/s of unique DRAM traffic)')
