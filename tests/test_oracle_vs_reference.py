"""Pin the travelling oracle against the reference's own X3 module run live (build container only: /root/reference
is absent on the GPU box, where these tests skip and the committed golden vectors take over)."""
import pytest
import torch

from oracle import e2_oracle as eo, ref_loader, synth

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason='/root/reference not present')


def _model(cfg, seed):
    tr = dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'],
              heads=cfg['heads'], dim_head=64, max_seq_len=cfg['max_seq_len'], if_text_modules=True,
              if_cross_attn=True, if_audio_conv=True, if_text_conv=True)
    m = ref_loader.build_reference_model(tr, num_channels=cfg['num_channels'])
    sd = synth.random_state_dict(**cfg, seed=seed)
    m.load_state_dict(sd, strict=False)
    return m, sd


def test_state_dict_names_and_shapes_match_reference_constructor():
    cfg = synth.TINY
    m, sd = _model(cfg, 0)
    ref = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith('video2roll_net')}
    assert ref == {k: tuple(v.shape) for k, v in sd.items()}


def test_default_init_conditioning_is_dead_and_rerandomised_is_live():
    cfg = synth.TINY
    sd0 = synth.random_state_dict(**cfg, seed=0, live_conditioning=False)
    bt = synth.batch([0], 40, dim_text=cfg['dim_text'], dim=cfg['dim'], d=cfg['num_channels'])
    arch = eo.Arch.from_state_dict(sd0)
    a = (bt['y0'], torch.tensor(0.3), None, bt['clip'], bt['frames'], bt['ctx'], bt['ctx_mask'])
    # CLIP and roll streams (and time) are dead at default init; the T5 cross-attention is not
    assert torch.equal(eo.pred_head(sd0, arch, *a), eo.pred_head(sd0, arch, *a, drop_clip=True, drop_frames=True))
    b = (bt['y0'], torch.tensor(0.9)) + a[2:]
    assert torch.equal(eo.pred_head(sd0, arch, *a), eo.pred_head(sd0, arch, *b))
    sd1 = synth.random_state_dict(**cfg, seed=0)
    assert not torch.allclose(eo.pred_head(sd1, arch, *a), eo.pred_head(sd1, arch, *a, drop_clip=True))
    assert not torch.allclose(eo.pred_head(sd1, arch, *a), eo.pred_head(sd1, arch, *b))


@pytest.mark.parametrize('apg', [False, True])
@pytest.mark.parametrize('lens', [[48, 48], [48, 31]])
def test_sample_bit_exact_vs_x3(apg, lens):
    cfg = synth.TINY
    m, sd = _model(cfg, 3)
    bt = synth.batch([5, 6], 48, lens=lens, nc_list=[7, 4], dim_text=cfg['dim_text'], dim=cfg['dim'],
                     d=cfg['num_channels'], live_frames=True)
    ref = ref_loader.reference_sample(m, y0=bt['y0'], clip=bt['clip'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'],
                                      frames_embed=bt['frames'], lens=bt['lens'], steps=5, cfg_strength=2.0,
                                      remove_parallel_component=apg)
    ours = eo.sample(sd, y0=bt['y0'], clip=bt['clip'], frames=bt['frames'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'],
                     lens=bt['lens'], steps=5, cfg_strength=2.0, remove_parallel_component=apg)
    assert ((ours - ref).norm() / ref.norm()).item() < 1e-6


@pytest.mark.parametrize('apg', [False, True])
def test_inpainting_bit_exact_vs_x3(apg):
    """lens < duration with E2TTS(if_cond_proj_in=True): cond_proj_in, step_cond, audio_drop_prompt and the final where
    (X3:2029-2035, 2224-2228, 2019-2020, 2259-2260)."""
    cfg = synth.TINY
    tr = dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'], heads=cfg['heads'], dim_head=64,
              max_seq_len=cfg['max_seq_len'], if_text_modules=True, if_cross_attn=True, if_audio_conv=True, if_text_conv=True)
    m = ref_loader.build_reference_model(tr, num_channels=cfg['num_channels'], if_cond_proj_in=True)
    sd = synth.random_state_dict(**cfg, seed=4, cond_proj_in=True)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith('video2roll_net') for k in missing)
    bt = synth.batch([5, 6], 48, lens=[48, 40], nc_list=[7, 4], dim_text=cfg['dim_text'], dim=cfg['dim'], d=cfg['num_channels'], live_frames=True)
    cond = torch.stack([synth.audio_condition(i, 48, cfg['num_channels']) for i in (5, 6)])
    cond_lens = torch.tensor([17, 40])
    ref = ref_loader.reference_sample(m, y0=bt['y0'], clip=bt['clip'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], frames_embed=bt['frames'],
                                      lens=bt['lens'], steps=4, cfg_strength=2.0, remove_parallel_component=apg, cond=cond, cond_lens=cond_lens,
                                      audio_drop_prompt=[False, True])
    ours = eo.sample(sd, y0=bt['y0'], clip=bt['clip'], frames=bt['frames'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], lens=bt['lens'], steps=4,
                     cfg_strength=2.0, remove_parallel_component=apg, cond=cond, cond_lens=cond_lens, audio_drop=[False, True])
    mask = eo.lens_to_mask(bt['lens'], 48)
    assert ((ours - ref)[mask].norm() / ref[mask].norm()).item() < 1e-6
