"""EnCodec decoder (csrc/encodec.cu) at the sampler's output shape: 10 s clips (750 latent frames -> 240 000 samples), random
facebook/encodec_24khz-shaped weights; per-kernel event profile and decoded audio-s/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200')]
import torch
from e2_tts_pytorch import _lib
from e2_tts_pytorch.encodec import EncodecDecoderB200
from oracle import encodec_oracle as eo
from oracle.make_golden_encodec import hf_decoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = 'cuda:0'
hf = hf_decoder(eo.DEFAULT, seed=0)
dec = EncodecDecoderB200({k: v.detach() for k, v in hf.state_dict().items()}, dev)
emb = torch.randn(B, 128, 750, device=dev)
for _ in range(2): out = dec(emb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): out = dec(emb)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f'decode {B} clips x 10 s: {ms:.1f} ms  = {B * 10 / (ms / 1e3):.0f} decoded audio-s/s   out {tuple(out.shape)}')
L = _lib.lib(); L.e2b_prof_enable(1); dec(emb); rows = _lib.profile_report(); L.e2b_prof_enable(0)
tot = sum(r['ms'] for r in rows)
for r in sorted(rows, key=lambda r: -r['ms'])[:14]:
    print(f"  {r['kind']:11s} rows={r['m']:8d} Co={r['n']:5d} K*Ci={r['k']:5d} x{r['count']:2d} {r['ms']:8.3f} ms {100*r['ms']/tot:5.1f} %  "
          f"{r['flops']*r['count']/(r['ms']*1e-3)/1e12:6.2f} TFLOP/s  {r['bytes']*r['count']/(r['ms']*1e-3)/1e9:7.1f} GB/s")
# the reference's own path on the same GPU: HuggingFace EncodecDecoder in torch eager (cuDNN convolutions and LSTM)
hfc = hf.to(dev)
with torch.no_grad():
    for _ in range(2): ref = hfc(emb)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3): ref = hfc(emb)
    e1.record(); torch.cuda.synchronize()
ms_ref = e0.elapsed_time(e1) / 3
err = ((out - ref).norm() / ref.norm()).item()
print(f'HuggingFace torch-eager decoder on the same GPU: {ms_ref:.1f} ms ({B * 10 / (ms_ref / 1e3):.0f} audio-s/s); rel-L2 between the two outputs {err:.1e}')
