"""EnCodec decoder kernels (csrc/encodec.cu) and the assembled decoder against the oracle (fp32: tolerance 1e-5 relative)."""
import os

import pytest
import torch
import torch.nn.functional as F

from gpu_util import DEV, L, kcheck, rel
from oracle import encodec_oracle as eo
from e2_tts_pytorch import _lib
from e2_tts_pytorch.encodec import ACCUMULATE, ELU_IN, REFLECT, EncodecDecoderB200

pytestmark = pytest.mark.gpu
sp, P = _lib.stream_ptr, _lib.ptr
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'encodec_tiny.pt')


@pytest.mark.parametrize('B,T,Ci,Co,K,flags', [(2, 37, 16, 32, 7, REFLECT), (1, 5, 8, 4, 3, ELU_IN | REFLECT), (3, 100, 128, 512, 7, REFLECT),
                                                (2, 300, 32, 1, 7, ELU_IN | REFLECT), (2, 64, 64, 40, 2, ELU_IN), (2, 33, 16, 16, 1, ACCUMULATE | ELU_IN),
                                                (1, 3, 8, 8, 7, REFLECT)])
def test_conv1d_cl(B, T, Ci, Co, K, flags):
    g = torch.Generator().manual_seed(T + Co)
    x = torch.randn(B, T, Ci, generator=g)
    w = torch.randn(Co, Ci, K, generator=g) / (Ci * K) ** 0.5
    bias = torch.randn(Co, generator=g)
    y0 = torch.randn(B, T, Co, generator=g)
    xin = F.elu(x) if flags & ELU_IN else x
    xp = xin.transpose(1, 2)
    xp = eo.pad_left_reflect(xp, K - 1) if flags & REFLECT else F.pad(xp, (K - 1, 0))
    ref = F.conv1d(xp, w, bias).transpose(1, 2)
    if flags & ACCUMULATE:
        ref = ref + y0
    raw = torch.full((B * T * Co + 2048,), 777.0, device=DEV)
    y = raw[1024:1024 + B * T * Co].view(B, T, Co)
    y.copy_(y0)
    wp = w.permute(2, 1, 0).contiguous().view(K * Ci, Co).to(DEV)
    xd, bd = x.to(DEV), bias.to(DEV)                      # (named: a temporary would be freed before the kernel runs)
    kcheck(L().e2b_conv1d_cl(P(xd), P(wp), P(bd), P(y), B, T, Ci, Co, K, Co, flags, sp()))
    torch.cuda.synchronize()
    assert (raw[:1024] == 777.0).all() and (raw[-1024:] == 777.0).all()
    assert rel(y.cpu(), ref) < 1e-5


@pytest.mark.parametrize('B,T,H', [(4, 9, 32), (8, 40, 512), (64, 6, 512)])
def test_lstm_layer(B, T, H):
    g = torch.Generator().manual_seed(B + H)
    x = torch.randn(B, H, T, generator=g)
    sd = {'l.lstm.weight_ih_l0': torch.randn(4 * H, H, generator=g) / H ** 0.5, 'l.lstm.weight_hh_l0': torch.randn(4 * H, H, generator=g) / H ** 0.5,
          'l.lstm.bias_ih_l0': torch.randn(4 * H, generator=g) * 0.1, 'l.lstm.bias_hh_l0': torch.randn(4 * H, generator=g) * 0.1}
    ref = eo.lstm(sd, 'l', x, 1).transpose(1, 2)                                      # [B, T, H], includes the skip
    xc = x.transpose(1, 2).contiguous().to(DEV)
    gx = (xc @ sd['l.lstm.weight_ih_l0'].t().to(DEV) + (sd['l.lstm.bias_ih_l0'] + sd['l.lstm.bias_hh_l0']).to(DEV)).contiguous()
    whh = sd['l.lstm.weight_hh_l0'].view(4, H // 4, 4, H).permute(1, 3, 2, 0).contiguous().to(DEV)
    out = torch.empty(B, T, H, device=DEV)
    hbuf = torch.empty(2, H, B, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    kcheck(L().e2b_lstm_layer(P(gx), P(whh), P(xc), P(out), P(hbuf), P(counter), B, T, H, sp()))
    assert rel(out.cpu(), ref) < 1e-5


def test_decoder_tiny_fixture_and_full_size():
    g = torch.load(GOLDEN, weights_only=False)
    dec = EncodecDecoderB200(g['sd'], DEV, upsampling_ratios=g['cfg']['upsampling_ratios'], num_lstm_layers=g['cfg']['num_lstm_layers'])
    out = dec(g['emb'].to(DEV))
    assert out.shape == g['out'].shape
    assert rel(out.cpu(), g['out']) < 1e-5                                            # the HuggingFace output itself
    # full-size decoder (facebook/encodec_24khz shapes, random weights): 3 clips of 0.8 s against the oracle
    from oracle.make_golden_encodec import hf_decoder
    hf = hf_decoder(eo.DEFAULT, seed=11)
    sd = {k: v.detach() for k, v in hf.state_dict().items()}
    emb = torch.randn(3, 128, 60)
    ref = eo.decode(sd, emb)
    out = EncodecDecoderB200(sd, DEV)(emb.to(DEV))
    assert out.shape == (3, 1, 60 * 320)
    e = rel(out.cpu(), ref)
    print(f'full-size EnCodec decoder: rel-L2 {e:.2e}')
    assert e < 1e-5


def test_batched_decode_of_padded_latents_equals_per_clip_decode():
    """The sampler's tail decodes the whole padded batch at once; by causality the valid prefix of every clip must equal the
    decode of that clip alone (what the reference's per-clip loop computes, X3:2277-2285)."""
    g = torch.load(GOLDEN, weights_only=False)
    dec = EncodecDecoderB200(g['sd'], DEV, upsampling_ratios=g['cfg']['upsampling_ratios'], num_lstm_layers=g['cfg']['num_lstm_layers'])
    emb = torch.randn(5, g['cfg']['hidden_size'], 23, device=DEV)
    lens = [23, 7, 16, 22, 9]                 # >= 7: shorter clips change the first conv's reflect padding (HF _pad1d)
    full = dec(emb)
    hop = full.shape[-1] // 23
    for i, n in enumerate(lens):
        one = dec(emb[i:i + 1, :, :n].contiguous())
        # (equal up to summation order: the LSTM kernel splits its dot products differently for different batch sizes)
        assert rel(one[0, 0], full[i, 0, :n * hop]) < 2e-6
        ref = eo.decode(g['sd'], emb[i:i + 1, :, :n].cpu(), g['cfg'])
        assert rel(one.cpu(), ref) < 1e-5
