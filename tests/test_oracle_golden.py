"""The travelling oracle (oracle/e2_oracle.py) against the committed golden vectors that oracle/make_golden.py
produced by running the reference's own X3 module (tests/golden/*.pt)."""
import os

import pytest
import torch

from oracle import e2_oracle as eo, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _load(name):
    g = torch.load(os.path.join(GOLD, name), weights_only=False)
    r = g['recipe']
    cfg = r['arch']
    sd = synth.random_state_dict(**cfg, seed=r['weight_seed'])
    bt = synth.batch(r['clips'], r['n'], lens=r['lens'], nc_list=r['nc_list'], dim_text=cfg['dim_text'],
                     dim=cfg['dim'], d=cfg['num_channels'], live_frames=r['live_frames'])
    return g, r, sd, bt


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_tiny_sample_and_single_passes_match_x3():
    g, r, sd, bt = _load('tiny_x3.pt')
    arch = eo.Arch.from_state_dict(sd)
    mask = eo.lens_to_mask(bt['lens'], r['n'])
    t = torch.tensor(r['t_single'])
    pc = eo.pred_head(sd, arch, bt['y0'], t, mask, bt['clip'], bt['frames'], bt['ctx'], bt['ctx_mask'])
    pn = eo.pred_head(sd, arch, bt['y0'], t, mask, bt['clip'], bt['frames'], bt['ctx'], bt['ctx_mask'],
                      drop_clip=True, drop_ctx=True)
    assert rel(pc, g['pred_cond']) < 1e-5 and rel(pn, g['pred_null']) < 1e-5
    assert rel(g['pred_cond'], g['pred_null']) > 1e-2          # conditioning is live
    kw = dict(y0=bt['y0'], clip=bt['clip'], frames=bt['frames'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'],
              lens=bt['lens'], steps=r['steps'], cfg_strength=r['cfg_strength'])
    assert rel(eo.sample(sd, **kw, remove_parallel_component=False), g['sample_cfg']) < 1e-5
    assert rel(eo.sample(sd, **kw, remove_parallel_component=True), g['sample_apg']) < 1e-5
    assert rel(g['sample_cfg'], g['sample_apg']) > 1e-3


@pytest.mark.timeout(600)
def test_shipped_single_pass_matches_x3():
    g, r, sd, bt = _load('shipped_x3.pt')
    arch = eo.Arch.from_state_dict(sd)
    assert (arch.depth, arch.dim, arch.dim_text, arch.dim_frames, arch.heads) == (12, 1024, 1280, 512, 16)
    mask = eo.lens_to_mask(bt['lens'], r['n'])
    with torch.no_grad():
        pc = eo.pred_head(sd, arch, bt['y0'], torch.tensor(r['t_single']), mask, bt['clip'], bt['frames'], bt['ctx'],
                          bt['ctx_mask'])
    assert rel(pc, g['pred_cond']) < 1e-5


def test_k_pass_guidance_reduces_to_cfg():
    g, r, sd, bt = _load('tiny_x3.pt')
    arch = eo.Arch.from_state_dict(sd)
    mask = eo.lens_to_mask(bt['lens'], r['n'])
    t = torch.tensor(0.2)
    a = (sd, arch, bt['y0'], t, mask, bt['clip'], bt['frames'], bt['ctx'], bt['ctx_mask'])
    v1 = eo.guided_velocity(*a, passes=(('null', 2.0),))
    pc, pn = eo.pred_head(*a), eo.pred_head(*a, drop_clip=True, drop_ctx=True)
    assert torch.allclose(v1, pc + (pc - pn) * 2.0)
    v2 = eo.guided_velocity(*a, passes=(('null', 1.0), ('drop_t5', 0.5), ('drop_roll', 0.25)))
    p5, pr = eo.pred_head(*a, drop_ctx=True), eo.pred_head(*a, drop_frames=True)
    assert torch.allclose(v2, pc + (pc - pn) + 0.5 * (pc - p5) + 0.25 * (pc - pr), atol=1e-5)
    assert torch.equal(eo.guided_velocity(*a, passes=(('null', 0.0),)), pc)


def test_flops_closed_form():
    assert abs(eo.flops_forward(750) / 1e9 - 1117.56) < 0.01
    assert abs(eo.flops_forward(2250) / 1e9 - 3680.82) < 0.01


def test_inpainting_matches_x3():
    """Audio-conditioned / in-painting mode: golden produced by the reference's own sample() with lens < duration
    and E2TTS(if_cond_proj_in=True) (oracle/make_golden.py::inpaint)."""
    g = torch.load(os.path.join(GOLD, 'tiny_x3_inpaint.pt'), weights_only=False)
    r = g['recipe']
    cfg = r['arch']
    sd = synth.random_state_dict(**cfg, seed=r['weight_seed'], cond_proj_in=True)
    bt = synth.batch(r['clips'], r['n'], lens=r['lens'], nc_list=r['nc_list'], dim_text=cfg['dim_text'], dim=cfg['dim'],
                     d=cfg['num_channels'], live_frames=r['live_frames'])
    cond = torch.stack([synth.audio_condition(i, r['n'], cfg['num_channels']) for i in r['clips']])
    mask = eo.lens_to_mask(bt['lens'], r['n'])
    kw = dict(y0=bt['y0'], clip=bt['clip'], frames=bt['frames'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], lens=bt['lens'],
              steps=r['steps'], cfg_strength=r['cfg_strength'], cond=cond, cond_lens=torch.tensor(r['cond_lens']), audio_drop=r['audio_drop'])
    for apg, key in ((False, 'sample_cfg'), (True, 'sample_apg')):
        out = eo.sample(sd, **kw, remove_parallel_component=apg)
        assert rel(out[mask], g[key][mask]) < 1e-5
        # conditioned frames come back verbatim (X3:2259-2260), the generated ones differ from an unconditioned run
        cm = eo.lens_to_mask(torch.tensor(r['cond_lens']), r['n'])
        assert torch.equal(out[cm], cond[cm])
    plain = eo.sample(sd, **{k: v for k, v in kw.items() if k not in ('cond', 'cond_lens', 'audio_drop')})
    gen = mask & ~eo.lens_to_mask(torch.tensor(r['cond_lens']), r['n'])
    assert rel(plain[gen], g['sample_cfg'][gen]) > 1e-3


@pytest.mark.parametrize('steps', [32, 64])
def test_long_trajectory_goldens_are_consistent(steps):
    """The 32- and 64-point shipped-architecture trajectories (oracle/make_golden.py::shipped_long): recorded grid = the sway
    grid of the oracle, last recorded state = the returned sample, states are distinct and finite.  (Running the oracle over
    the whole trajectory takes minutes per clip; the GPU suite compares the CUDA path against these files, and the oracle is
    pinned to X3 bit for bit at 2-6 updates by the tests above.)"""
    path = os.path.join(GOLD, f'shipped_x3_s{steps}.pt')
    if not os.path.exists(path):
        pytest.skip('fixture not generated')
    g = torch.load(path, weights_only=False)
    r = g['recipe']
    assert r['steps'] == steps and r['n'] == 750 and r['cfg_strength'] == 2.0
    assert torch.equal(g['grid'], eo.sway_grid(steps))
    assert g['updates'][-1] == steps - 1 and len(g['updates']) == g['states'].shape[0]
    assert torch.equal(g['states'][-1], g['sample_cfg'])
    assert torch.isfinite(g['states']).all()
    for a, b in zip(g['states'][:-1], g['states'][1:]):
        assert rel(a, b) > 1e-2
