"""Piano-roll front end on the GPU: e2b_frame_windows / e2b_roll_expand and the drop-in E2TTS.encode_frames
against the oracle restatement of X3:1525-1555 and the fixture the reference's own encode_frames produced.  The window
gather is a copy (bit-exact); the roll goes through the stand-in net on the GPU, so it is compared to 1e-5."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import DEV, L, kcheck
from oracle import frames_oracle as fo, synth
from oracle.make_golden_frames import ENCODE_FRAMES_CASES
from e2_tts_pytorch import _lib
from test_gpu_4_path import build_model

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'frames.npz'))


@pytest.mark.parametrize('b,t,w,h', [(2, 9, 100, 900), (1, 1, 100, 900), (3, 4, 8, 12), (1, 2, 100, 900)])
def test_frame_windows_bit_exact(b, t, w, h):
    x = torch.rand(b, 1, t, w, h, generator=torch.Generator().manual_seed(b * 100 + t))
    ref = fo.frame_windows(x)
    xd = x.to(DEV)
    out = torch.full((b * t + 1, 5, w, h), -7.0, device=DEV)                          # one guard frame behind the output
    kcheck(L().e2b_frame_windows(_lib.ptr(xd), _lib.ptr(out), b, t, w * h, 5, _lib.stream_ptr()))
    assert torch.equal(out[:-1].cpu(), ref)
    assert (out[-1] == -7.0).all()


@pytest.mark.parametrize('b,t,l', [(2, 9, 30), (1, 5, 11), (1, 6, 18), (3, 251, 750)])
def test_roll_expand_vs_torch(b, t, l):
    logits = torch.randn(b * t, 51, generator=torch.Generator().manual_seed(t)) * 3
    ref = torch.sigmoid(logits).reshape(b, t, 1, 51).repeat(1, 1, 3, 1).reshape(b, 3 * t, 51)
    ref = ref[:, :l] if 3 * t >= l else torch.cat((ref, torch.zeros(b, l - 3 * t, 51)), 1)
    out = torch.full((b * l + 1, 51), -7.0, device=DEV)
    kcheck(L().e2b_roll_expand(_lib.ptr(logits.to(DEV)), _lib.ptr(out), b, t, l, 51, 3, _lib.stream_ptr()))
    got = out[:-1].reshape(b, l, 51).cpu()
    assert (got - ref).abs().max().item() < 1e-6
    if 3 * t < l:
        assert not got[:, 3 * t:].any()
    assert (out[-1] == -7.0).all()


@pytest.mark.parametrize('k', range(len(ENCODE_FRAMES_CASES)))
def test_encode_frames_dropin_vs_reference_fixture(k):
    seed, b, t, l = ENCODE_FRAMES_CASES[k]
    m, _ = build_model(synth.TINY)
    m.video2roll_net = synth.StandInRollNet().to(DEV)
    x = torch.stack([synth.grey_frames(seed + 100 * i, t)[..., 0] for i in range(b)])[:, None]
    roll = m.encode_frames(x.to(DEV), l)
    ref = torch.from_numpy(GOLD[f'roll{k}'])
    assert tuple(roll.shape) == (b, l, 51)
    assert (roll.cpu() - ref).abs().max().item() < 1e-5


def test_sample_takes_the_frame_stack():
    """sample(frames=Float[b,1,t,100,900]) runs encode_frames -> proj_frames on the GPU (X3:2170) and matches passing the roll."""
    cfg = synth.TINY
    m, _ = build_model(cfg)
    m.video2roll_net = synth.StandInRollNet().to(DEV)
    n, b = 40, 2
    bt = {k: v.to(DEV) for k, v in synth.batch([0, 1], n, dim_text=cfg['dim_text'], dim=cfg['dim'], d=cfg['num_channels']).items()}
    x = torch.stack([synth.grey_frames(50 + i, n // 3 + 1)[..., 0] for i in range(b)])[:, None].to(DEV)
    kw = dict(text=bt['clip'], lens=bt['lens'], duration=bt['lens'], steps=3, cfg_strength=2.0, remove_parallel_component=False,
              return_raw_output=True, context=bt['ctx'], context_mask=bt['ctx_mask'], noise=bt['y0'])
    a = m.sample(torch.zeros_like(bt['y0']), frames=x, **kw)
    roll = fo.encode_frames(x.cpu(), n, synth.StandInRollNet())
    c = m.sample(torch.zeros_like(bt['y0']), frames=roll.to(DEV), **kw)
    z = m.sample(torch.zeros_like(bt['y0']), frames=None, **kw)
    assert ((a - c).norm() / c.norm()).item() < 1e-3
    assert ((a - z).norm() / z.norm()).item() > 1e-3        # the roll stream is live
