"""Generates tests/golden/encodec_tiny.pt with the HuggingFace EncodecDecoder (transformers): a reduced configuration with
random weights, its input and its output.  The full-size decoder is compared live (transformers travels with the image)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TINY = dict(hidden_size=16, num_filters=8, upsampling_ratios=(4, 2), kernel_size=7, last_kernel_size=7, residual_kernel_size=3,
            num_lstm_layers=2, audio_channels=1, compress=2)


def hf_decoder(cfg, seed=0):
    from transformers import EncodecConfig, EncodecModel
    torch.manual_seed(seed)
    c = EncodecConfig(hidden_size=cfg['hidden_size'], num_filters=cfg['num_filters'], upsampling_ratios=list(cfg['upsampling_ratios']),
                      kernel_size=cfg['kernel_size'], last_kernel_size=cfg['last_kernel_size'], residual_kernel_size=cfg['residual_kernel_size'],
                      num_lstm_layers=cfg['num_lstm_layers'], audio_channels=cfg['audio_channels'], compress=cfg['compress'],
                      codebook_dim=cfg['hidden_size'])
    dec = EncodecModel(c).decoder.eval()
    with torch.no_grad():                               # weight-norm g is initialised to ||v||: scramble it so it matters
        for k, p in dec.named_parameters():
            if k.endswith('original0'):
                p.mul_(torch.rand_like(p) + 0.5)
            if k.endswith('.bias') or 'bias_' in k:
                p.normal_(0, 0.1)
    return dec


def main():
    dec = hf_decoder(TINY)
    torch.manual_seed(1)
    emb = torch.randn(3, TINY['hidden_size'], 9)
    with torch.no_grad():
        out = dec(emb)
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    torch.save(dict(cfg=TINY, sd=sd, emb=emb, out=out), os.path.join(ROOT, 'tests', 'golden', 'encodec_tiny.pt'))
    print('wrote', tuple(out.shape))


if __name__ == '__main__':
    main()
