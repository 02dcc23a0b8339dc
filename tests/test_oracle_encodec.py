"""EnCodec decoder oracle against the committed HuggingFace fixture and against the HuggingFace module live."""
import os

import pytest
import torch

from oracle import encodec_oracle as eo

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'encodec_tiny.pt')


def test_oracle_matches_huggingface_fixture():
    g = torch.load(GOLDEN, weights_only=False)
    out = eo.decode(g['sd'], g['emb'], g['cfg'])
    assert out.shape == g['out'].shape
    assert (out - g['out']).abs().max().item() < 1e-6


def test_building_blocks():
    # reflect padding of a short signal is zero-extended first (HF _pad1d)
    x = torch.tensor([[[1.0, 2.0]]])
    assert eo.pad_left_reflect(x, 3).tolist() == [[[0.0, 0.0, 2.0, 1.0, 2.0]]]
    assert eo.pad_left_reflect(torch.arange(5.0).view(1, 1, 5), 2).tolist() == [[[2.0, 1.0, 0.0, 1.0, 2.0, 3.0, 4.0]]]
    # transposed conv: length T * stride after the causal trim
    sd = {'c.conv.parametrizations.weight.original0': torch.ones(3, 1, 1), 'c.conv.parametrizations.weight.original1': torch.randn(3, 2, 8),
          'c.conv.bias': torch.zeros(2)}
    assert eo.conv_transpose1d(sd, 'c', torch.randn(1, 3, 5), 4).shape == (1, 2, 20)


def test_oracle_matches_huggingface_live():
    pytest.importorskip('transformers')
    from oracle.make_golden_encodec import hf_decoder
    for cfg, T in ((eo.DEFAULT, 30), (dict(eo.DEFAULT, num_filters=8, upsampling_ratios=(4, 2), hidden_size=16), 5)):
        dec = hf_decoder(cfg, seed=5)
        emb = torch.randn(2, cfg['hidden_size'], T)
        with torch.no_grad():
            ref = dec(emb)
        out = eo.decode({k: v.detach() for k, v in dec.state_dict().items()}, emb, cfg)
        assert (out - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
