"""End-to-end parity of the CUDA hot path (through the drop-in Python surface -> C-ABI) against the oracle and the
golden vectors produced by the reference's own X3 module.  Tolerance: relative L2 <= 1e-2 (bf16 mode), the bound
BASELINE.json's north_star states."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import DEV, rel
from oracle import e2_oracle as eo, synth
from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS, Transformer

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = 1e-2


def build_model(cfg, seed=0):
    tr = dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'], heads=cfg['heads'],
              dim_head=64, max_seq_len=cfg['max_seq_len'], if_text_modules=True, if_cross_attn=True, if_audio_conv=True,
              if_text_conv=True)
    m = E2TTS(duration_predictor=None, transformer=tr, tokenizer='char_utf8', audiocond_drop_prob=1.1, cond_drop_prob=-0.1,
              prompt_drop_prob=-0.1, if_cond_proj_in=False, if_embed_text=False, if_text_encoder2=False, if_clip_encoder=False,
              num_channels=cfg['num_channels'], sampling_rate=24000)
    sd = synth.random_state_dict(**cfg, seed=seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m.to(DEV), sd


def load_gold(name):
    g = torch.load(os.path.join(GOLD, name), weights_only=False)
    r = g['recipe']
    cfg = r['arch']
    bt = synth.batch(r['clips'], r['n'], lens=r['lens'], nc_list=r['nc_list'], dim_text=cfg['dim_text'], dim=cfg['dim'],
                     d=cfg['num_channels'], live_frames=r['live_frames'])
    return g, r, cfg, bt


def dev(bt):
    return {k: v.to(DEV) for k, v in bt.items()}


def valid_rel(a, b, lens):
    mask = eo.lens_to_mask(torch.as_tensor(lens), a.shape[-2]).to(a.device)
    return rel(a[mask], b.to(a.device)[mask])


def test_tiny_velocity_vs_x3_golden():
    g, r, cfg, bt = load_gold('tiny_x3.pt')
    m, _ = build_model(cfg, r['weight_seed'])
    d = dev(bt)
    pred = m.velocity(d['y0'], r['t_single'], clip=d['clip'], context=d['ctx'], context_mask=d['ctx_mask'], roll=d['frames'],
                      lens=r['lens'], passes=('null',))
    e0, e1 = valid_rel(pred[0], g['pred_cond'], r['lens']), valid_rel(pred[1], g['pred_null'], r['lens'])
    print(f'tiny velocity rel-L2: cond {e0:.3e} null {e1:.3e}')
    assert e0 < TOL and e1 < TOL


@pytest.mark.parametrize('apg', [False, True])
def test_tiny_sample_vs_x3_golden(apg):
    g, r, cfg, bt = load_gold('tiny_x3.pt')
    m, _ = build_model(cfg, r['weight_seed'])
    d = dev(bt)
    out = m.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=d['lens'], duration=d['lens'], steps=r['steps'],
                   cfg_strength=r['cfg_strength'], remove_parallel_component=apg, sway_sampling=True, return_raw_output=True,
                   context=d['ctx'], context_mask=d['ctx_mask'], frames=d['frames'], noise=d['y0'])
    ref = g['sample_apg' if apg else 'sample_cfg']
    e = valid_rel(out, ref, r['lens'])
    print(f'tiny sample apg={apg}: rel-L2 {e:.3e}; launches {m._last_launches}')
    assert e < TOL
    assert m._last_launches > 0


def test_tiny_kpass_guidance_vs_oracle():
    g, r, cfg, bt = load_gold('tiny_x3.pt')
    m, sd = build_model(cfg, r['weight_seed'])
    d = dev(bt)
    passes = [('null', 1.0), ('drop_t5', 0.5), ('drop_roll', 0.75), ('drop_clip', -0.25)]
    out = m.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=d['lens'], duration=d['lens'], steps=4, guidance=passes,
                   remove_parallel_component=False, return_raw_output=True, context=d['ctx'], context_mask=d['ctx_mask'],
                   frames=d['frames'], noise=d['y0'])
    ref = eo.sample(sd, y0=bt['y0'], clip=bt['clip'], frames=bt['frames'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], lens=bt['lens'],
                    steps=4, passes=passes)
    e = valid_rel(out, ref, r['lens'])
    print(f'tiny K-pass sample: rel-L2 {e:.3e}')
    assert e < TOL


def test_transformer_forward_dropin_vs_oracle():
    cfg = synth.TINY
    sd = synth.random_state_dict(**cfg, seed=2)
    tr = Transformer(dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'], depth=cfg['depth'], heads=cfg['heads'],
                     dim_head=64, max_seq_len=cfg['max_seq_len'], if_text_conv=True)
    tr.load_state_dict({k[len('transformer.'):]: v for k, v in sd.items() if k.startswith('transformer.')})
    tr = tr.to(DEV)
    b, n, nc = 2, 45, 6
    g = torch.Generator().manual_seed(0)
    x = torch.randn(b, n, cfg['dim'], generator=g)
    text = torch.randn(b, n, cfg['dim_text'], generator=g)
    fr = torch.randn(b, n, cfg['dim_frames'], generator=g)
    ctx = torch.randn(b, nc, cfg['dim'], generator=g)
    times = torch.tensor([0.2, 0.7])
    lens = torch.tensor([45, 30])
    mask = eo.lens_to_mask(lens, n)
    cmask = eo.lens_to_mask(torch.tensor([6, 4]), nc)
    arch = eo.Arch.from_state_dict(sd)
    with torch.no_grad():
        ref = eo.transformer_forward(sd, arch, x, times, mask, text, fr, ctx, cmask)
    out = tr(x.to(DEV), times=times.to(DEV), mask=mask.to(DEV), text_embed=text.to(DEV), frames_embed=fr.to(DEV),
             context=ctx.to(DEV), context_mask=cmask.to(DEV))
    e = valid_rel(out, ref, lens.tolist())
    print(f'Transformer.forward drop-in: rel-L2 {e:.3e}')
    assert e < TOL


def test_shipped_arch_vs_x3_golden():
    """Full-size architecture (776 M parameters), one 10 s clip: single velocity passes and a 3-point sample."""
    g, r, cfg, bt = load_gold('shipped_x3.pt')
    m, _ = build_model(cfg, r['weight_seed'])
    d = dev(bt)
    pred = m.velocity(d['y0'], r['t_single'], clip=d['clip'], context=d['ctx'], context_mask=d['ctx_mask'], roll=None, lens=r['lens'],
                      passes=('null',))
    e0, e1 = rel(pred[0], g['pred_cond'].to(DEV)), rel(pred[1], g['pred_null'].to(DEV))
    out = m.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=d['lens'], duration=d['lens'], steps=r['steps'],
                   cfg_strength=r['cfg_strength'], remove_parallel_component=False, return_raw_output=True, context=d['ctx'],
                   context_mask=d['ctx_mask'], noise=d['y0'])
    e2 = rel(out, g['sample_cfg'].to(DEV))
    print(f'shipped arch rel-L2: cond {e0:.3e} null {e1:.3e} sample {e2:.3e}')
    assert e0 < TOL and e1 < TOL and e2 < TOL


def test_size_independent_properties_batch64():
    """Full C2 batch shape: per-clip results must not depend on batch composition (each clip's ODE is independent) and
    clips with identical conditions must give identical latents."""
    cfg = synth.SHIPPED
    m, _ = build_model(cfg, 0)
    n, B = 750, 8
    bt = dev(synth.batch(list(range(B)), n))
    kw = dict(steps=3, cfg_strength=2.0, remove_parallel_component=False, return_raw_output=True)
    full = m.sample(torch.zeros_like(bt['y0']), text=bt['clip'], lens=bt['lens'], duration=bt['lens'], context=bt['ctx'],
                    context_mask=bt['ctx_mask'], noise=bt['y0'], **kw)
    sub = m.sample(torch.zeros_like(bt['y0'][2:5]), text=bt['clip'][2:5], lens=bt['lens'][2:5], duration=bt['lens'][2:5],
                   context=bt['ctx'][2:5], context_mask=bt['ctx_mask'][2:5], noise=bt['y0'][2:5], **kw)
    assert torch.isfinite(full).all()
    assert torch.equal(full[2:5], sub)
    assert not torch.equal(full[0], full[1])


@pytest.mark.parametrize('apg', [False, True])
def test_graph_replay_is_bit_identical_to_eager(apg):
    """e2b_sample runs eagerly on a first call, captures a CUDA graph of the whole step loop on the second call with the same
    signature and replays it afterwards: all three must agree bit for bit, also when the conditions change between calls."""
    g, r, cfg, bt = load_gold('tiny_x3.pt')
    m, _ = build_model(cfg, r['weight_seed'])
    d = dev(bt)

    def run(y0, clip):
        return m.sample(torch.zeros_like(y0), text=clip, lens=d['lens'], duration=d['lens'], steps=r['steps'],
                        cfg_strength=r['cfg_strength'], remove_parallel_component=apg, sway_sampling=True, return_raw_output=True,
                        context=d['ctx'], context_mask=d['ctx_mask'], frames=d['frames'], noise=y0)

    eager = run(d['y0'], d['clip'])
    l_eager = m._last_launches
    captured = run(d['y0'].clone(), d['clip'])
    replayed = run(d['y0'].clone(), d['clip'])
    assert torch.equal(eager, captured) and torch.equal(eager, replayed)
    assert m._last_launches == l_eager                     # replay reports the launches of the captured sequence
    # new data through the same graph
    y1, clip1 = torch.randn_like(d['y0']), torch.randn_like(d['clip'])
    via_graph = run(y1, clip1)
    # a different signature (other step count) goes back to the eager path; then compare the new data eagerly
    m.sample(torch.zeros_like(y1), text=clip1, lens=d['lens'], duration=d['lens'], steps=r['steps'] + 1, cfg_strength=r['cfg_strength'],
             remove_parallel_component=apg, return_raw_output=True, context=d['ctx'], context_mask=d['ctx_mask'], frames=d['frames'], noise=y1)
    eager1 = run(y1, clip1)                                # first sighting after the signature changed: eager again
    assert torch.equal(via_graph, eager1)
    assert not torch.equal(via_graph, eager)


@pytest.mark.parametrize('gold', ['tiny_x3.pt', 'shipped_x3.pt'])
def test_branch_overlap_is_bit_identical_to_one_stream(gold, monkeypatch):
    """Small batches run the text / frames branches on their own streams beside the audio stream (own scratch buffers, event
    ordering around the pre-update bf16 copies).  Same kernels, same data: the latents must equal the one-stream schedule bit for
    bit -- eagerly, while the graph is captured, and replayed."""
    g, r, cfg, bt = load_gold(gold)
    d = dev(bt)
    outs = {}
    for rows in ('0', '1000000'):
        monkeypatch.setenv('E2B_OVERLAP_ROWS', rows)       # read when the workspace of a shape is allocated
        m, _ = build_model(cfg, r['weight_seed'])
        run = lambda: m.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=d['lens'], duration=d['lens'], steps=r['steps'],
                               cfg_strength=r['cfg_strength'], remove_parallel_component=False, sway_sampling=True, return_raw_output=True,
                               context=d['ctx'], context_mask=d['ctx_mask'], frames=d['frames'], noise=d['y0'].clone())
        outs[rows] = [run(), run(), run()]                 # eager, captured, replayed
        del m
    for a in outs['0'] + outs['1000000']:
        assert torch.equal(a, outs['0'][0])


def test_graph_replay_follows_new_lengths_and_context():
    """Clip lengths, context lengths and pass data live in device buffers that set_conditions rewrites: a replayed graph must use
    the new ones (nothing length-dependent may be baked into a captured launch)."""
    g, r, cfg, bt = load_gold('tiny_x3.pt')
    m, _ = build_model(cfg, r['weight_seed'])
    d = dev(bt)
    n = d['y0'].shape[1]

    def run(lens, ctx_mask, steps=r['steps']):
        return m.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=lens, duration=lens, steps=steps, cfg_strength=r['cfg_strength'],
                        remove_parallel_component=False, sway_sampling=True, return_raw_output=True, context=d['ctx'], context_mask=ctx_mask,
                        frames=d['frames'], noise=d['y0'])

    full = torch.full_like(d['lens'], n)
    run(full, d['ctx_mask'])
    run(full, d['ctx_mask'])                                  # captured
    short = torch.clamp(d['lens'] - 3, min=1)
    short[0] = n                                              # keep the padded length
    cm = d['ctx_mask'].clone()
    cm[:, -1] = False
    cm[:, 0] = True
    via_graph = run(short, cm)                                # replayed with new lengths / context mask
    run(short, cm, steps=r['steps'] + 1)                      # other signature: drops back to eager ...
    eager = run(short, cm)                                    # ... and so does the first call after it
    mask = eo.lens_to_mask(short.cpu(), n)
    assert torch.equal(via_graph.cpu()[mask], eager.cpu()[mask])
