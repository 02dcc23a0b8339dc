"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by running the reference's own X3 module.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Weights and inputs are NOT stored: they are regenerated from oracle/synth.py by seed / clip index, so the
fixtures hold only the reference outputs (a few hundred KB).  Each file records the recipe that produced it.
"""
from __future__ import annotations

import os
import time

import torch

from . import ref_loader, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _ref_model(cfg, seed, cond_proj_in=False):
    tr = dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'],
              heads=cfg['heads'], dim_head=cfg['dim_head'], max_seq_len=cfg['max_seq_len'], if_text_modules=True,
              if_cross_attn=True, if_audio_conv=True, if_text_conv=True)
    m = ref_loader.build_reference_model(tr, num_channels=cfg['num_channels'], if_cond_proj_in=cond_proj_in)
    sd = synth.random_state_dict(**cfg, seed=seed, cond_proj_in=cond_proj_in)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith('video2roll_net') for k in missing), (missing, unexpected)
    return m


def _velocity(m, bt, t, drop):
    """One reference transformer_with_pred_head call (X3:1993) with both drops on or off."""
    import contextlib, io
    m.encode_text = lambda prompt: (bt['ctx'].clone(), bt['ctx_mask'].clone())
    b, n, _ = bt['y0'].shape
    mask = torch.arange(n)[None, :] < bt['lens'][:, None]
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        return m.transformer_with_pred_head(
            bt['y0'].clone(), None, times=torch.tensor(t), mask=mask, text=bt['clip'].clone(),
            frames_embed=bt['frames'].clone(), prompt=['x'] * b, video_drop_prompt=[False] * b,
            drop_audio_cond=drop, drop_text_cond=drop, drop_text_prompt=drop)


def tiny():
    cfg, seed = synth.TINY, 0
    m = _ref_model(cfg, seed)
    rec = dict(arch=cfg, weight_seed=seed, clips=[0, 1, 2], n=50, lens=[50, 37, 44], nc_list=[8, 5, 6],
               live_frames=True, steps=6, cfg_strength=2.0, t_single=0.37)
    bt = synth.batch(rec['clips'], rec['n'], lens=rec['lens'], nc_list=rec['nc_list'], dim_text=cfg['dim_text'],
                     dim=cfg['dim'], d=cfg['num_channels'], live_frames=True)
    out = dict(recipe=rec)
    kw = dict(y0=bt['y0'], clip=bt['clip'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], frames_embed=bt['frames'],
              lens=bt['lens'], steps=rec['steps'], cfg_strength=rec['cfg_strength'])
    out['sample_cfg'] = ref_loader.reference_sample(m, remove_parallel_component=False, **kw)
    out['sample_apg'] = ref_loader.reference_sample(m, remove_parallel_component=True, **kw)
    out['pred_cond'] = _velocity(m, bt, rec['t_single'], False)
    out['pred_null'] = _velocity(m, bt, rec['t_single'], True)
    out = {k: (v.clone().contiguous() if torch.is_tensor(v) else v) for k, v in out.items()}
    torch.save(out, os.path.join(OUT, 'tiny_x3.pt'))
    print('tiny_x3.pt', {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})


def inpaint():
    """Audio-conditioned / in-painting mode: E2TTS(if_cond_proj_in=True), lens < duration, audiocond_snr None.
    Clip 1 has its audio condition dropped through audio_drop_prompt (X3:2019-2020); clip 2 is conditioned over its whole
    length (the whole-batch switch at X3:2224 only looks at clip 0)."""
    cfg, seed = synth.TINY, 3
    m = _ref_model(cfg, seed, cond_proj_in=True)
    rec = dict(arch=cfg, weight_seed=seed, clips=[0, 1, 2], n=50, lens=[50, 37, 44], cond_lens=[20, 30, 44], audio_drop=[False, True, False],
               nc_list=[8, 5, 6], live_frames=True, steps=6, cfg_strength=2.0)
    bt = synth.batch(rec['clips'], rec['n'], lens=rec['lens'], nc_list=rec['nc_list'], dim_text=cfg['dim_text'],
                     dim=cfg['dim'], d=cfg['num_channels'], live_frames=True)
    cond = torch.stack([synth.audio_condition(i, rec['n'], cfg['num_channels']) for i in rec['clips']])
    kw = dict(y0=bt['y0'], clip=bt['clip'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], frames_embed=bt['frames'], lens=bt['lens'],
              steps=rec['steps'], cfg_strength=rec['cfg_strength'], cond=cond, cond_lens=torch.tensor(rec['cond_lens']),
              audio_drop_prompt=rec['audio_drop'])
    out = dict(recipe=rec)
    out['sample_cfg'] = ref_loader.reference_sample(m, remove_parallel_component=False, **kw)
    out['sample_apg'] = ref_loader.reference_sample(m, remove_parallel_component=True, **kw)
    out = {k: (v.clone().contiguous() if torch.is_tensor(v) else v) for k, v in out.items()}
    torch.save(out, os.path.join(OUT, 'tiny_x3_inpaint.pt'))
    print('tiny_x3_inpaint.pt', {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})


def shipped():
    cfg, seed = synth.SHIPPED, 0
    torch.set_num_threads(os.cpu_count())
    m = _ref_model(cfg, seed)
    rec = dict(arch=cfg, weight_seed=seed, clips=[0], n=750, lens=[750], nc_list=[8], live_frames=False, steps=3,
               cfg_strength=2.0, t_single=0.3)
    bt = synth.batch(rec['clips'], rec['n'], nc=8)
    out = dict(recipe=rec)
    t0 = time.time()
    out['pred_cond'] = _velocity(m, bt, rec['t_single'], False)
    out['pred_null'] = _velocity(m, bt, rec['t_single'], True)
    out['sample_cfg'] = ref_loader.reference_sample(
        m, y0=bt['y0'], clip=bt['clip'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], lens=bt['lens'], steps=rec['steps'],
        cfg_strength=rec['cfg_strength'], remove_parallel_component=False)
    print(f'shipped: 6 reference forwards in {time.time() - t0:.1f}s')
    out = {k: (v.clone().contiguous() if torch.is_tensor(v) else v) for k, v in out.items()}
    torch.save(out, os.path.join(OUT, 'shipped_x3.pt'))
    print('shipped_x3.pt', {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})




def shipped_long(steps, every=8):
    """Shipped 776 M architecture, clip 0, n = 750, CFG 2.0, sway grid: the whole `steps`-point trajectory of the reference's
    own sample() (C1 of BASELINE.json at steps=32; steps=64 is the CLI's setting, src/inference_v2a.py:183).  Stores the final
    latent and the ODE state after every `every` Euler updates (the reference's odeint returns the full trajectory, X3:2255)."""
    cfg, seed = synth.SHIPPED, 0
    torch.set_num_threads(os.cpu_count())
    m = _ref_model(cfg, seed)
    x3 = ref_loader.load_x3()
    rec = dict(arch=cfg, weight_seed=seed, clips=[0], n=750, lens=[750], nc_list=[8], live_frames=False, steps=steps,
               cfg_strength=2.0, every=every)
    bt = synth.batch(rec['clips'], rec['n'], nc=8)
    captured = []
    real_odeint = x3.odeint

    def recording_odeint(fn, y0, t, **kw):
        traj = real_odeint(fn, y0, t, **kw)
        captured.append((traj, t.clone()))
        return traj

    x3.odeint = recording_odeint
    t0 = time.time()
    try:
        final = ref_loader.reference_sample(
            m, y0=bt['y0'], clip=bt['clip'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], lens=bt['lens'], steps=steps,
            cfg_strength=rec['cfg_strength'], remove_parallel_component=False)
    finally:
        x3.odeint = real_odeint
    traj, grid = captured[0]
    assert traj.shape[0] == steps and torch.equal(traj[-1], final)
    idx = sorted(set(list(range(every, steps - 1, every)) + [steps - 1]))      # number of Euler updates applied
    out = dict(recipe=rec, sample_cfg=final.clone().contiguous(), grid=grid, updates=idx,
               states=torch.stack([traj[i] for i in idx]).clone().contiguous())
    print(f'shipped steps={steps}: {2 * (steps - 1)} reference forwards in {time.time() - t0:.1f}s; states after updates {idx}')
    torch.save(out, os.path.join(OUT, f'shipped_x3_s{steps}.pt'))


if __name__ == '__main__':
    import sys
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ['tiny', 'inpaint', 'shipped', 'long32', 'long64']
    for w in which:
        {'tiny': tiny, 'shipped': shipped, 'inpaint': inpaint, 'long32': lambda: shipped_long(32), 'long64': lambda: shipped_long(64)}[w]()
