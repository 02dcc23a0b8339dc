"""Achieved HBM GB/s of the HBM-bound kernels of the path at their C2 sizes (north_star: "achieved HBM GB/s for the elementwise and
STFT kernels"): STFT+mel front end, guided Euler update (CFG and APG), piano-roll front-end kernels.  CUDA events, inputs larger than
L2 or flushed between launches; algorithmic bytes as DESIGN.md states them."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
import torch
import torchaudio
from e2_tts_pytorch import _lib
from e2_tts_pytorch.e2_tts_crossatt3 import MelSpec
from gpu_util import L, kcheck, DEV

sp, P = _lib.stream_ptr, _lib.ptr
flush = torch.empty(64 * 1024 * 1024, device=DEV)          # 256 MB > L2


def timed(fn, reps=5):
    fn()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        fn()
        e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def report(name, us, nbytes, extra=''):
    print(f'{name:44s} {us:9.1f} us   {nbytes / us / 1e3:8.1f} GB/s of {nbytes / 1e6:8.1f} MB algorithmic  {extra}', flush=True)


# ---- STFT + mel (A13): 64 clips x 10 s at 24 kHz -> [64, 100, 938]
B, nw = 64, 240000
wav = torch.rand(B, nw, device=DEV) - 0.5
mel = MelSpec().to(DEV)
out = mel(wav)
ref = torchaudio.transforms.MelSpectrogram(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=100, power=1, center=True,
                                           normalized=False, norm=None).to(DEV)(wav).clamp(min=1e-5).log()
err = ((out - ref).norm() / ref.norm()).item()
T = out.shape[-1]
us = timed(lambda: mel(wav))
flops = B * T / 2 * 5 * 1024 * 10
report('melspec 64 x 10 s', us, 4.0 * B * (nw + 100 * T), f'(rel err {err:.1e}; ~{flops / us / 1e6:.2f} TFLOP/s of FFT arithmetic)')
us_ref = timed(lambda: ref_mod(wav)) if False else None
tm = torchaudio.transforms.MelSpectrogram(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=100, power=1, center=True,
                                          normalized=False, norm=None).to(DEV)
report('  torchaudio (cuFFT + matmul) same input', timed(lambda: tm(wav).clamp(min=1e-5).log()), 4.0 * B * (nw + 100 * T))

# ---- guided Euler (A3) at C2: P = 2 passes, 64 clips x 750 x 128
for Pn, apg, label in ((2, 0, 'guided_euler CFG P=2'), (2, 1, 'guided_euler APG P=2'), (4, 0, 'guided_euler K-pass P=4')):
    Bc, per = 64, 750 * 128
    y = torch.randn(Bc, per, device=DEV)
    pred = torch.randn(Pn, Bc, per, device=DEV)
    yb = torch.empty(Pn, Bc, per, device=DEV, dtype=torch.bfloat16)
    scratch = torch.zeros(2 * Bc, device=DEV, dtype=torch.float64)
    w = (C.c_float * (Pn - 1))(*([2.0] + [0.5] * (Pn - 2)))
    fn = lambda: kcheck(L().e2b_guided_euler_launch(P(y), P(pred), Pn, Bc, per, w, 0.03, apg, 0.0, P(scratch), P(yb), Pn, sp()))
    nbytes = 4.0 * Bc * per * (Pn + 2) + 2.0 * Bc * per * Pn + (4.0 * Bc * per * 2 if apg else 0)
    report(label, timed(fn), nbytes)

# ---- piano-roll front end (N3): 16 clips x 251 frames of 100 x 900 -> 5-frame windows; roll expansion to 750 frames
b, t = 4, 251
x = torch.rand(b, 1, t, 100, 900, device=DEV)
win = torch.empty(b * t, 5, 100, 900, device=DEV)
report('frame_windows 4 x 251 x (100x900)', timed(lambda: kcheck(L().e2b_frame_windows(P(x), P(win), b, t, 90000, 5, sp()))), 4.0 * b * t * 90000 * 6)
logits = torch.randn(64 * 251, 51, device=DEV)
roll = torch.empty(64, 750, 51, device=DEV)
report('roll_expand 64 x 251 -> 750', timed(lambda: kcheck(L().e2b_roll_expand(P(logits), P(roll), 64, 251, 750, 51, 3, sp()))), 4.0 * 64 * 51 * (251 + 750))
