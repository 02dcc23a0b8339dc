"""Per-kernel event profile of one Euler update at a small batch (default: the reference CLI's single 10 s clip)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200')]
import torch
import bench
from oracle import synth
from e2_tts_pytorch import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if os.environ.get('EW8_MAX_K'):      # experiment knob: largest K that takes the 8-epilogue-warp GEMM configuration
    import ctypes
    ctypes.c_int.in_dll(_lib.lib(), 'e2b_gemm_ew8_max_k').value = int(os.environ['EW8_MAX_K'])
dev = torch.device('cuda', 0)
model, _ = bench.shipped_model(dev)
bt = {k: v.to(dev) for k, v in synth.batch(list(range(B)), 750).items()}
def run(steps):
    return model.sample(torch.zeros_like(bt['y0']), text=bt['clip'], lens=bt['lens'], duration=bt['lens'], context=bt['ctx'],
                        context_mask=bt['ctx_mask'], noise=bt['y0'], steps=steps, cfg_strength=2.0, remove_parallel_component=False,
                        return_raw_output=True)
run(3); run(3)
torch.cuda.synchronize()
L = _lib.lib()
L.e2b_prof_enable(1)
run(2)
rows = _lib.profile_report()
L.e2b_prof_enable(0)
tot = sum(r['ms'] for r in rows)
kinds = {}
for r in rows:
    k = kinds.setdefault(r['kind'], [0.0, 0])
    k[0] += r['ms']; k[1] += r['count']
print(f'B={B}: one Euler update = {tot:.3f} ms over {sum(r["count"] for r in rows)} launches')
for k, (ms, n) in sorted(kinds.items(), key=lambda kv: -kv[1][0]):
    print(f'  {k:14s} {ms:7.3f} ms  {100 * ms / tot:5.1f} %  {n:4d} launches  {1e3 * ms / n:7.1f} us each')
for r in sorted(rows, key=lambda r: -r['ms'])[:int(os.environ.get('TOP', 12))]:
    print(f"  {r['kind']:12s} m={r['m']:6d} n={r['n']:5d} k={r['k']:5d} x{r['count']:3d} {1e3*r['ms']/r['count']:7.1f} us  {r['flops']*r['count']/(r['ms']*1e-3)/1e12:7.1f} TF/s")
