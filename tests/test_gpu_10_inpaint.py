"""Audio-conditioned / in-painting sampling (lens < duration, E2TTS(if_cond_proj_in=True)) through the drop-in
sample() -> e2b_set_audio_cond -> e2b_sample, against the golden the reference's own sample() produced
(tests/golden/tiny_x3_inpaint.pt, oracle/make_golden.py::inpaint).  Tolerances as everywhere: rel-L2 <= 1e-2 (bf16), <= 1e-4 (fp32)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import DEV, rel
from oracle import e2_oracle as eo, synth
from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _model(cfg, seed, cond_proj_in=True):
    tr = dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'], heads=cfg['heads'], dim_head=64,
              max_seq_len=cfg['max_seq_len'], if_text_modules=True, if_cross_attn=True, if_audio_conv=True, if_text_conv=True)
    m = E2TTS(duration_predictor=None, transformer=tr, tokenizer='char_utf8', audiocond_drop_prob=1.1, cond_drop_prob=-0.1, prompt_drop_prob=-0.1,
              if_cond_proj_in=cond_proj_in, if_embed_text=False, if_text_encoder2=False, if_clip_encoder=False, num_channels=cfg['num_channels'],
              sampling_rate=24000)
    sd = synth.random_state_dict(**cfg, seed=seed, cond_proj_in=cond_proj_in)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m.to(DEV), sd


def _inputs():
    g = torch.load(os.path.join(GOLD, 'tiny_x3_inpaint.pt'), weights_only=False)
    r = g['recipe']
    cfg = r['arch']
    bt = synth.batch(r['clips'], r['n'], lens=r['lens'], nc_list=r['nc_list'], dim_text=cfg['dim_text'], dim=cfg['dim'], d=cfg['num_channels'],
                     live_frames=r['live_frames'])
    cond = torch.stack([synth.audio_condition(i, r['n'], cfg['num_channels']) for i in r['clips']])
    return g, r, cfg, {k: v.to(DEV) for k, v in bt.items()}, cond.to(DEV)


def _run(m, r, d, cond, apg, **kw):
    # reference call shape: cond = the clip's latent, lens = conditioned frames, duration = total frames (X3:2196-2216)
    return m.sample(cond, text=d['clip'], lens=torch.tensor(r['cond_lens'], device=DEV), duration=d['lens'], steps=r['steps'],
                    cfg_strength=r['cfg_strength'], remove_parallel_component=apg, sway_sampling=True, return_raw_output=True,
                    context=d['ctx'], context_mask=d['ctx_mask'], frames=d['frames'], noise=d['y0'], audio_drop_prompt=r['audio_drop'], **kw)


@pytest.mark.parametrize('precision', ['bf16', 'fp32'])
@pytest.mark.parametrize('apg', [False, True])
def test_inpainting_vs_x3_golden(apg, precision):
    g, r, cfg, d, cond = _inputs()
    m, _ = _model(cfg, r['weight_seed'])
    m.precision = precision
    out = _run(m, r, d, cond, apg)
    ref = g['sample_apg' if apg else 'sample_cfg'].to(DEV)
    mask = eo.lens_to_mask(d['lens'], r['n'])
    cm = eo.lens_to_mask(torch.tensor(r['cond_lens'], device=DEV), r['n'])
    gen = mask & ~cm
    e_all, e_gen = rel(out[mask], ref[mask]), rel(out[gen], ref[gen])
    print(f'in-painting apg={apg} {precision}: rel-L2 valid rows {e_all:.3e}, generated rows only {e_gen:.3e}')
    tol = 1e-2 if precision == 'bf16' else 1e-4
    assert e_all < tol and e_gen < tol
    assert torch.equal(out[cm], cond[cm])                    # conditioned frames are returned verbatim (X3:2259-2260)


def test_inpainting_graph_replay_and_mode_switch():
    """The captured step loop must carry the in-painting select, and a later plain call (lens == duration) on the same engine must
    not see the stale condition."""
    g, r, cfg, d, cond = _inputs()
    m, sd = _model(cfg, r['weight_seed'])
    a = _run(m, r, d, cond, False)
    b = _run(m, r, d, cond, False)          # captured
    c = _run(m, r, d, cond, False)          # replayed
    assert torch.equal(a, b) and torch.equal(a, c)
    cond2 = torch.randn_like(cond)
    via_graph = _run(m, r, d, cond2, False)
    cm = eo.lens_to_mask(torch.tensor(r['cond_lens'], device=DEV), r['n'])
    assert torch.equal(via_graph[cm], cond2[cm]) and not torch.equal(via_graph, a)
    plain = m.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=d['lens'], duration=d['lens'], steps=r['steps'], cfg_strength=r['cfg_strength'],
                     remove_parallel_component=False, return_raw_output=True, context=d['ctx'], context_mask=d['ctx_mask'], frames=d['frames'],
                     noise=d['y0'])
    bt = {k: v.cpu() for k, v in d.items()}
    ref = eo.sample(sd, y0=bt['y0'], clip=bt['clip'], frames=bt['frames'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], lens=bt['lens'],
                    steps=r['steps'], cfg_strength=r['cfg_strength'])
    mask = eo.lens_to_mask(bt['lens'], r['n'])
    assert rel(plain.cpu()[mask], ref[mask]) < 1e-2


def test_inpainting_needs_cond_proj_in_and_no_snr():
    g, r, cfg, d, cond = _inputs()
    m, _ = _model(cfg, r['weight_seed'], cond_proj_in=False)
    with pytest.raises(TypeError, match='NoneType'):        # the reference calls self.cond_proj_in = None here (X3:2034)
        _run(m, r, d, cond, False)
    m2, _ = _model(cfg, r['weight_seed'])
    m2.audiocond_snr = (5.0, 10.0)
    with pytest.raises(IndexError):                          # X3:2123 indexes [b,n,d] with the [b,n,1] cond_mask
        _run(m2, r, d, cond, False)


def test_sample_rejects_mismatched_shapes():
    """Shapes reach libe2b as raw pointers: mismatches must raise on the host like the reference's first matmul would."""
    g, r, cfg, d, cond = _inputs()
    m, _ = _model(cfg, r['weight_seed'])
    kw = dict(lens=d['lens'], duration=d['lens'], steps=2, cfg_strength=2.0, return_raw_output=True, context=d['ctx'], context_mask=d['ctx_mask'])
    z = torch.zeros_like(d['y0'])
    with pytest.raises(ValueError, match='text'):
        m.sample(z, text=d['clip'][:, :, :-1], **kw)
    with pytest.raises(ValueError, match='text'):
        m.sample(z, text=d['clip'][:1], **kw)
    with pytest.raises(ValueError, match='text has'):
        m.sample(z, text=d['clip'][:, :10], **kw)
    with pytest.raises(ValueError, match='context'):
        m.sample(z, text=d['clip'], **dict(kw, context=d['ctx'][..., :-1]))
    with pytest.raises(ValueError, match='context_mask'):
        m.sample(z, text=d['clip'], **dict(kw, context_mask=d['ctx_mask'][:, :-1]))
    with pytest.raises(ValueError, match='noise'):
        m.sample(z, text=d['clip'], noise=d['y0'][:, :-1], **kw)
    with pytest.raises(ValueError, match='piano roll'):
        m.sample(z, text=d['clip'], frames=d['frames'][..., :-1], **kw)
