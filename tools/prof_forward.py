"""Small driver for ncu: one CFM Euler update (2-pass CFG) of the shipped architecture at a reduced batch."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200')]
import torch

import bench
from oracle import synth

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--frames', type=int, default=750)
ap.add_argument('--updates', type=int, default=1)
a = ap.parse_args()
dev = torch.device('cuda', 0)
model, _ = bench.shipped_model(dev)
bt = {k: v.to(dev) for k, v in synth.batch(list(range(a.batch)), a.frames).items()}
out = model.sample(torch.zeros_like(bt['y0']), text=bt['clip'], lens=bt['lens'], duration=bt['lens'], context=bt['ctx'],
                   context_mask=bt['ctx_mask'], noise=bt['y0'], steps=a.updates + 1, cfg_strength=2.0, remove_parallel_component=False,
                   return_raw_output=True)
torch.cuda.synchronize()
print('ok', float(out.abs().mean()))
