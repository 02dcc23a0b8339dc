#!/usr/bin/env python
"""bench.py -- CFM sampling throughput (generated audio-seconds per second) of the B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C2|C3|C4|C5]

One "step" = one full CFM sampling pass (E2TTS.sample: sway grid, 32 grid points = 31 Euler updates, 2-pass CFG) over one
batch of synthetic 10 s clips -- BASELINE.json configs[1] (C2: 64 clips per GPU, bf16, shipped 776 M-parameter
architecture, random-init weights).  N > 1 (torchrun): every rank samples its own 64 clips (weak scaling, sharded by clip,
per-clip seeds keyed by global clip index) and the outputs are all-gathered over NCCL at the end of each step.

The other BASELINE.json configurations are selectable (the default and the driver's run stay C2):
  --config C3   512 clips in total, sharded by clip over the ranks (strong scaling: 512 / 256 / 128 / 64 clips per GPU at
                N = 1 / 2 / 4 / 8), sampled in sub-batches of 64
  --config C4   V2P: 16 clips of 30 s (n = 2250, N = 2282: long-sequence attention) with the piano-roll stream live
  --config C5   K-pass guidance (3 guided passes per ODE step: null, drop_t5, drop_clip), default 64 grid points, 20 s clips;
                sweep with --sample-steps 16..64 and --frames 375..2250

Prints ONE JSON line (rank 0).  `value` is measured with the conditions already resident in HBM; `e2e` is the same metric
through the public API with pinned HOST inputs (H2D of CLIP/T5/noise + D2H of the latents inside the timed region).
`roofline` comes from a separate profiled pass (CUDA events around every launch, csrc/prof.cu); `cpu_baseline` is the
oracle (oracle/e2_oracle.py, a port of the reference's PyTorch-eager sampler) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

FRAME_RATE = 75.0
METRIC = 'generated audio-sec/sec (CFM sampling)'
UNIT = 'audio-s/s'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], burst=p['bf16_tflops'], sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']), src='measured')
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


CONFIGS = {
    'C2': dict(batch=64, frames=750, sample_steps=32, guidance=None, live_roll=False, total=None,
               name='C2: V2A CFM sampling, {B} synthetic {sec:.0f} s clips per GPU (CLIP+T5 cond)'),
    'C3': dict(batch=64, frames=750, sample_steps=32, guidance=None, live_roll=False, total=512,
               name='C3: V2A CFM sampling, 512 synthetic {sec:.0f} s clips sharded by clip over the ranks, sub-batches of {B}'),
    'C4': dict(batch=16, frames=2250, sample_steps=32, guidance=None, live_roll=True, total=None,
               name='C4: V2P piano generation, {B} synthetic {sec:.0f} s two-hand clips per GPU (CLIP+T5+live piano roll)'),
    'C5': dict(batch=16, frames=1500, sample_steps=64, guidance=[('null', 2.0), ('drop_t5', 0.5), ('drop_clip', 0.5)], live_roll=False, total=None,
               name='C5: V2A with K-pass guidance (3 guided passes per ODE step), {B} synthetic {sec:.0f} s clips per GPU'),
}


def shipped_model(device):
    import synthetic as synth
    from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS
    cfg = synth.SHIPPED
    tr = dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'], heads=cfg['heads'],
              dim_head=64, max_seq_len=cfg['max_seq_len'], if_text_modules=True, if_cross_attn=True, if_audio_conv=True,
              if_text_conv=True)
    m = E2TTS(duration_predictor=None, transformer=tr, tokenizer='phoneme_zh', audiocond_drop_prob=1.1, cond_drop_prob=-0.1,
              prompt_drop_prob=-0.1, if_cond_proj_in=False, if_embed_text=False, if_text_encoder2=False, if_clip_encoder=False,
              num_channels=cfg['num_channels'], sampling_rate=24000)
    sd = synth.random_state_dict(**cfg, seed=0)
    m.load_state_dict(sd, strict=True)
    return m.to(device), sd


_REF_MODEL = {}


def cpu_reference_kind():
    """'reference' when the reference's own module is importable on this box (its X3 file staged under baseline/_ref by
    oracle/stage_reference.py, or /root/reference itself), else 'port' (oracle/e2_oracle.py)."""
    from oracle import ref_loader
    return 'reference' if ref_loader.reference_available() else 'port'


def cpu_port_seconds_per_update(sd, n, threads, updates=1, clip_index=0, live_roll=False, passes=None):
    """The reference's eager fp32 sampler on the host cores: seconds per Euler update (one forward per guidance pass), B=1.
    Runs the reference's OWN E2TTS.sample (X3 imported through oracle/ref_loader.py) when it is available on this box and the
    call is plain 2-pass CFG; the oracle port of it otherwise (same arithmetic, bit-exact against X3 in the tests)."""
    from oracle import e2_oracle as eo, ref_loader
    import synthetic as synth
    torch.set_num_threads(threads)
    bt = synth.batch([clip_index], n, live_frames=live_roll)
    t0 = time.perf_counter()
    if passes is None and ref_loader.reference_available():
        m = _REF_MODEL.get('m')
        if m is None:
            m = ref_loader.build_reference_model(dict(ref_loader.SHIPPED_TRANSFORMER))
            m.load_state_dict(sd, strict=False)
            _REF_MODEL['m'] = m
            t0 = time.perf_counter()
        ref_loader.reference_sample(m, y0=bt['y0'], clip=bt['clip'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], lens=bt['lens'],
                                    frames_embed=bt['frames'] if live_roll else None, steps=updates + 1, cfg_strength=2.0,
                                    remove_parallel_component=False)
    else:
        arch = eo.Arch.from_state_dict(sd)
        eo.sample(sd, y0=bt['y0'], clip=bt['clip'], frames=bt['frames'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], lens=bt['lens'],
                  steps=updates + 1, cfg_strength=2.0, passes=passes, arch=arch)
    return (time.perf_counter() - t0) / updates


def run_reference(args, rank):
    """--impl reference: the reference's CPU sampler (oracle port; the reference is Python and cannot travel) on this box's
    host cores.  Each step = one Euler update (cond + null forward) of ONE 10 s clip; value extrapolates to the full
    32-point grid of the same workload."""
    if rank != 0:
        return
    import synthetic as synth
    threads = os.cpu_count() or 1
    sd = synth.random_state_dict(**synth.SHIPPED, seed=0)
    n, grid_points = args.frames, args.sample_steps
    passes = CONFIGS[args.config]['guidance']
    live = CONFIGS[args.config]['live_roll']
    kind = cpu_reference_kind() if passes is None else 'port'
    fwd = 2 if passes is None else 1 + len(passes)
    for _ in range(args.warmup):
        cpu_port_seconds_per_update(sd, n, threads, live_roll=live, passes=passes)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_port_seconds_per_update(sd, n, threads, clip_index=i, live_roll=live, passes=passes)
    per_update = (time.perf_counter() - t0) / max(args.steps, 1)
    value = (n / FRAME_RATE) / (per_update * (grid_points - 1))
    what = "the reference's own E2TTS.sample (X3 through oracle/ref_loader.py)" if kind == 'reference' else \
        'oracle port of the reference eager sampler (oracle/e2_oracle.py)'
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=per_update * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                impl='reference',
                config=dict(workload=f'{args.config}: CFM sampling, {n}-frame ({n / FRAME_RATE:.0f} s) clips, {grid_points} grid points, '
                                     f'{fwd} forwards per update, shipped 776M arch; CPU arm ({what}): 1 clip, one Euler update per step, '
                                     f'extrapolated x{grid_points - 1}'),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind=kind,
                                  sample=f'B=1, n={n}: {args.steps} Euler updates ({fwd} forwards each) timed, scaled to {grid_points - 1} updates'),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def _load_traffic():
    out = {}
    for name in ('r01_ncu_traffic.json', 'r02_ncu_traffic.json'):      # later rounds override
        try:
            with open(os.path.join(ROOT, 'profiles', name)) as f:
                out.update(json.load(f))
        except OSError:
            pass
    return out


NCU_TRAFFIC = _load_traffic()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='C2', choices=sorted(CONFIGS), help='BASELINE.json configuration (default C2)')
    ap.add_argument('--batch', type=int, default=None, help='clips per GPU and sample() call')
    ap.add_argument('--frames', type=int, default=None, help='latent frames per clip (750 = 10 s)')
    ap.add_argument('--sample-steps', type=int, default=None, help='ODE grid points (32 => 31 Euler updates)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel event profile to this file')
    args = ap.parse_args()
    conf = CONFIGS[args.config]
    for k in ('batch', 'frames', 'sample_steps'):
        if getattr(args, k) is None:
            setattr(args, k, conf[k])

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank)

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU path; use --impl reference for the CPU arm)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    import synthetic as synth
    from e2_tts_pytorch import _lib
    from e2_tts_pytorch.sharding import shard_clips
    B, n, S = args.batch, args.frames, args.sample_steps
    model, sd = shipped_model(dev)
    if conf['total']:                                            # strong scaling: a fixed clip list sharded by clip
        mine = list(shard_clips(conf['total'], rank, world))
        total_clips = conf['total']
    else:                                                        # weak scaling: B clips per GPU
        mine = [rank * B + i for i in range(B)]
        total_clips = world * B
    chunks = [mine[i:i + B] for i in range(0, len(mine), B)]     # one sample() call per chunk of <= B clips
    guidance = conf['guidance']
    P = 2 if guidance is None else 1 + len(guidance)
    keys = ('y0', 'clip', 'ctx') + (('frames',) if conf['live_roll'] else ())
    host = [synth.batch(c, n, live_frames=conf['live_roll']) for c in chunks]     # fp32 host tensors, keyed by global clip index
    pin = [{k: v.pin_memory() for k, v in h.items() if k in keys} for h in host]
    ctx_mask = [h['ctx_mask'].to(dev) for h in host]
    lens = [h['lens'].to(dev) for h in host]
    res = [{k: v.to(dev) for k, v in p.items()} for p in pin]
    out_host = [torch.empty(len(c), n, 128).pin_memory() for c in chunks]
    local_out = torch.empty(len(mine), n, 128, device=dev)
    gathered = torch.empty(world * len(mine), n, 128, device=dev) if world > 1 else None
    kw = dict(steps=S, remove_parallel_component=False, sway_sampling=True, return_raw_output=True)
    kw.update(dict(cfg_strength=2.0) if guidance is None else dict(guidance=guidance))

    compute_events = []                                          # per step: events around this rank's own sampling (collective excluded)

    def run_chunks(data, to_host):
        at = 0
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        for i, d in enumerate(data):
            if to_host:
                d = {k: v.to(dev, non_blocking=True) for k, v in d.items()}
            out = model.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=lens[i], duration=lens[i], context=d['ctx'],
                               context_mask=ctx_mask[i], noise=d['y0'], frames=d.get('frames'), **kw)
            local_out[at:at + out.shape[0]] = out
            at += out.shape[0]
            if to_host:
                out_host[i].copy_(out, non_blocking=True)
        ev[1].record()
        compute_events.append(ev)
        if world > 1:
            dist.all_gather_into_tensor(gathered, local_out)     # the path's only collective (equal shards: 512 % world == 0)
        return local_out

    def step_resident():
        return run_chunks(res, False)

    def step_e2e():
        return run_chunks(pin, True)

    rank_ms = [None]

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        mine_ms = e0.elapsed_time(e1)
        ms = torch.tensor([mine_ms], device=dev)
        if world > 1:
            every = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(every, ms)
            rank_ms[0] = [round(float(v.item()), 3) for v in every]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(args.warmup):
        step_resident()
    eng = model.engine()
    l0 = eng.launch_count()
    del compute_events[:]
    with ClockSampler(local) as cs:
        ms = timed(step_resident, args.steps)
    launches = eng.launch_count() - l0
    clocks = cs.summary()
    own_ms = sum(a.elapsed_time(b) for a, b in compute_events)   # this rank's sampling alone, without waiting for the other ranks
    per_rank = None
    if world > 1:                                                # attribution of the scaling loss: every rank's own time and clock
        mine = torch.tensor([clocks['sm_mhz'] or 0.0, own_ms], device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank = dict(ms_timed_region=rank_ms[0], ms_own_sampling=[round(float(v[1].item()), 3) for v in every],
                        sm_mhz_median=[float(v[0].item()) for v in every])
    audio_s = total_clips * (n / FRAME_RATE) * args.steps
    value = audio_s / (ms / 1e3)

    step_e2e()
    k_e2e = max(1, min(args.steps, 2))
    ms_e2e = timed(step_e2e, k_e2e)
    e2e_value = total_clips * (n / FRAME_RATE) * k_e2e / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for p in pin for v in p.values())
    d2h = sum(o.numel() * o.element_size() for o in out_host)

    # ---- profiled pass (rank 0): one Euler update with CUDA events around every launch --------------------------------------
    pk = peaks()
    roof, fwd = None, None
    if rank == 0:
        L = _lib.lib()
        torch.cuda.synchronize()
        L.e2b_prof_enable(1)
        r0 = res[0]
        model.sample(torch.zeros_like(r0['y0']), text=r0['clip'], lens=lens[0], duration=lens[0], context=r0['ctx'], context_mask=ctx_mask[0],
                     noise=r0['y0'], frames=r0.get('frames'), **dict(kw, steps=2))
        rows = _lib.profile_report()
        L.e2b_prof_enable(0)
        total_ms = sum(r['ms'] for r in rows)
        tensor_rows = [r for r in rows if r['kind'].startswith('gemm') or r['kind'] == 'attention']
        top = max(rows, key=lambda r: r['ms'])
        per_launch_ms = top['ms'] / top['count']
        achieved = top['flops'] / (per_launch_ms * 1e-3) / 1e12
        def roof_of(r):
            ms_l = r['ms'] / r['count']
            ach = r['flops'] / (ms_l * 1e-3) / 1e12
            key = f"{r['kind']} M={r['m']} N={r['n']} K={r['k']}"
            tr = NCU_TRAFFIC.get(key)             # dram bytes per launch from the committed ncu --set full capture (or None)
            return dict(bound='tensor', kernel=key, achieved=ach, peak=pk['burst'], unit='TFLOP/s', frac=ach / pk['burst'],
                        peak_source=f"{pk['src']} burst bf16 (MEASURED_PEAKS.json)", launches_per_update=r['count'], ms_per_launch=ms_l,
                        share_of_step=r['ms'] / total_ms, algorithmic_bytes=r['bytes'],
                        traffic=tr['dram_bytes_per_launch'] if tr else None, traffic_source=tr['source'] if tr else None)
        roof = roof_of(top)
        if top['kind'] == 'attention':
            # the attention kernel's binding unit is the MUFU pipe (one ex2 per logit at 16 / clk / SM = 1024 clk per 128x128 tile
            # against 512 clk of MMA for head dim 64), so its tensor-pipe fraction cannot exceed ~0.5 (profiles/r01_pipe_microbench.txt)
            roof['note'] = 'MUFU-bound kernel (ex2 16/clk/SM): tensor fraction ceiling ~0.5 at head dim 64'
            roof['top_gemm'] = roof_of(max((r for r in rows if r['kind'].startswith('gemm')), key=lambda r: r['ms']))
        flops_update = eng.flops_per_forward()                   # both passes, as executed
        tf_all = flops_update / (total_ms * 1e-3) / 1e12
        by_kind = {}
        for r in rows:
            k = by_kind.setdefault(r['kind'], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            k['ms'] += r['ms']; k['flops'] += r['flops'] * r['count']; k['bytes'] += r['bytes'] * r['count']; k['launches'] += r['count']
        fwd = dict(flops_per_update=flops_update, ms_per_update_profiled=total_ms, tflops=tf_all, frac_of_sustained=tf_all / pk['sustained'],
                   tensor_kernel_share=sum(r['ms'] for r in tensor_rows) / total_ms,
                   schedule='one stream: the library switches the three-stream branch overlap off while the event profiler is on, so these '
                            'are unperturbed per-kernel times (their sum is a few percent above the overlapped step)',
                   by_kind={k: dict(ms=round(v['ms'], 3), share=round(v['ms'] / total_ms, 4),
                                    tflops=round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 1) if v['flops'] else None,
                                    gbs=round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 1), launches=v['launches'])
                            for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1]['ms'])})
        if args.profile_out:
            with open(args.profile_out, 'w') as f:
                json.dump(dict(rows=rows, summary=fwd, roofline=roof), f, indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ck = cpu_reference_kind() if guidance is None else 'port'
        cpu_port_seconds_per_update(sd, n, threads, live_roll=conf['live_roll'], passes=guidance)     # warm-up (page-in, thread pool)
        nup = 2 if n <= 750 else 1
        per = cpu_port_seconds_per_update(sd, n, threads, updates=nup, live_roll=conf['live_roll'], passes=guidance)
        cpu = dict(value=(n / FRAME_RATE) / (per * (S - 1)), unit=UNIT, cores=threads, kind=ck,
                   sample=("the reference's own eager fp32 E2TTS.sample (X3 via oracle/ref_loader.py)" if ck == 'reference' else
                           'oracle (torch fp32 port of the reference eager sampler)') +
                          f', B=1, n={n}: {nup} of {S - 1} Euler updates ({nup * P} forwards, {per:.2f} s/update) timed and scaled')

    if rank == 0:
        # e2b_forward_flops() is per sample() call at the prepared shape (all passes of one chunk), per rank
        flops_per_gpu = eng.flops_per_forward() * (S - 1) * len(chunks) * args.steps
        wl = conf['name'].format(B=B, sec=n / FRAME_RATE) + \
            (f', {S} grid points ({S - 1} Euler updates) with ' + ('2-pass CFG 2.0' if guidance is None else
                                                                  f'{P}-pass guidance {guidance}') +
             ', shipped 776M 3-stream arch, random-init weights')
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='strong' if conf['total'] else 'weak', vs_baseline=None,
                    dtype='bf16', data='synthetic',
                    config=dict(workload=wl, config=args.config, clips_per_gpu=len(mine), clips_per_call=B, clips_total=total_clips, frames=n,
                                sample_steps=S, guidance_passes=P,
                                parallelism=f'shard-by-clip x{world}' if world > 1 else 'single GPU',
                                l2_policy='working set >> L2 (GBs of activations streamed per forward); no flush needed'),
                    tflops_executed=flops_per_gpu / (ms * 1e-3) / 1e12,
                    tensor_util_vs_sustained=flops_per_gpu / (ms * 1e-3) / 1e12 / pk['sustained'],
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=launches, clocks=clocks, per_rank=per_rank, roofline=roof, forward=fwd, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
