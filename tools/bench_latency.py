"""Latency of one sample() call for small batches (the reference CLI's shape: one 10 s clip, 32 grid points, 2-pass CFG),
eager launches vs the captured CUDA graph of the step loop (E2B_GRAPH=0 disables the graph)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200')]
import torch
import bench
from oracle import synth
dev = torch.device('cuda:0')
model, _ = bench.shipped_model(dev)
n = 750
for B in (1, 2, 4, 8):
    host = synth.batch(list(range(B)), n)
    res = {k: v.to(dev) for k, v in host.items()}
    lens = torch.full((B,), n, device=dev, dtype=torch.long)
    ctx_mask = res['ctx_mask']
    def run():
        return model.sample(torch.zeros_like(res['y0']), text=res['clip'], lens=lens, duration=lens, context=res['ctx'], context_mask=ctx_mask,
                            noise=res['y0'], steps=32, cfg_strength=2.0, sway_sampling=True, remove_parallel_component=False, return_raw_output=True)
    for _ in range(3): run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 5
    for _ in range(K): run()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / K * 1e3
    print(f'B={B}: {ms:8.1f} ms per sample() call  ({B * n / 75 / (ms / 1e3):6.1f} audio-s/s)  graph={os.environ.get("E2B_GRAPH", "1")}')
