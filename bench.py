#!/usr/bin/env python
"""bench.py -- CFM sampling throughput (generated audio-seconds per second) of the B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one full CFM sampling pass (E2TTS.sample: sway grid, 32 grid points = 31 Euler updates, 2-pass CFG) over one
batch of synthetic 10 s clips -- BASELINE.json configs[1] (C2: 64 clips per GPU, bf16, shipped 776 M-parameter
architecture, random-init weights).  N > 1 (torchrun): every rank samples its own 64 clips (weak scaling, sharded by clip,
per-clip seeds keyed by global clip index) and the outputs are all-gathered over NCCL at the end of each step.

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


def shipped_model(device):
    from oracle import synth
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


def cpu_port_seconds_per_update(sd, n, threads, updates=1, clip_index=0):
    """Oracle (port of the reference's eager sampler) on the host cores: seconds per Euler update (2 forwards), B=1."""
    from oracle import e2_oracle as eo, synth
    torch.set_num_threads(threads)
    bt = synth.batch([clip_index], n)
    arch = eo.Arch.from_state_dict(sd)
    t0 = time.perf_counter()
    eo.sample(sd, y0=bt['y0'], clip=bt['clip'], frames=bt['frames'], ctx=bt['ctx'], ctx_mask=bt['ctx_mask'], lens=bt['lens'],
              steps=updates + 1, cfg_strength=2.0, arch=arch)
    return (time.perf_counter() - t0) / updates


def run_reference(args, rank):
    """--impl reference: the reference's CPU sampler (oracle port; the reference is Python and cannot travel) on this box's
    host cores.  Each step = one Euler update (cond + null forward) of ONE 10 s clip; value extrapolates to the full
    32-point grid of the same workload."""
    if rank != 0:
        return
    from oracle import synth
    threads = os.cpu_count() or 1
    sd = synth.random_state_dict(**synth.SHIPPED, seed=0)
    n, grid_points = args.frames, args.sample_steps
    for _ in range(args.warmup):
        cpu_port_seconds_per_update(sd, n, threads)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_port_seconds_per_update(sd, n, threads, clip_index=i)
    per_update = (time.perf_counter() - t0) / max(args.steps, 1)
    value = (n / FRAME_RATE) / (per_update * (grid_points - 1))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=per_update * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                impl='reference',
                config=dict(workload=f'V2A CFM sampling, {n}-frame (10 s) clips, {grid_points} grid points, CFG 2.0, shipped 776M arch; '
                                     f'reference arm: 1 clip, one Euler update per step, extrapolated x{grid_points - 1}'),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind='port',
                                  sample=f'B=1, n={n}: {args.steps} Euler updates (2 forwards each) timed, scaled to {grid_points - 1} updates'),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def _load_traffic():
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')) as f:
            return json.load(f)
    except OSError:
        return {}


NCU_TRAFFIC = _load_traffic()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='clips per GPU')
    ap.add_argument('--frames', type=int, default=750, help='latent frames per clip (750 = 10 s)')
    ap.add_argument('--sample-steps', type=int, default=32, help='ODE grid points (32 => 31 Euler updates)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel event profile to this file')
    args = ap.parse_args()

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

    from oracle import synth
    from e2_tts_pytorch import _lib
    B, n, S = args.batch, args.frames, args.sample_steps
    model, sd = shipped_model(dev)
    clips = [rank * B + i for i in range(B)]
    host = synth.batch(clips, n)                                 # fp32 host tensors, keyed by global clip index
    pin = {k: v.pin_memory() for k, v in host.items() if k in ('y0', 'clip', 'ctx')}
    ctx_mask = host['ctx_mask'].to(dev)
    lens = host['lens'].to(dev)
    res = {k: v.to(dev) for k, v in pin.items()}
    out_host = torch.empty(B, n, 128).pin_memory()
    gathered = torch.empty(world * B, n, 128, device=dev) if world > 1 else None
    kw = dict(steps=S, cfg_strength=2.0, remove_parallel_component=False, sway_sampling=True, return_raw_output=True)

    def step_resident():
        out = model.sample(torch.zeros_like(res['y0']), text=res['clip'], lens=lens, duration=lens, context=res['ctx'],
                           context_mask=ctx_mask, noise=res['y0'], **kw)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        return out

    def step_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in pin.items()}
        out = model.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=lens, duration=lens, context=d['ctx'],
                           context_mask=ctx_mask, noise=d['y0'], **kw)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        out_host.copy_(out, non_blocking=True)
        return out

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
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(args.warmup):
        step_resident()
    eng = model.engine()
    l0 = eng.launch_count()
    with ClockSampler(local) as cs:
        ms = timed(step_resident, args.steps)
    launches = eng.launch_count() - l0
    clocks = cs.summary()
    audio_s = world * B * (n / FRAME_RATE) * args.steps
    value = audio_s / (ms / 1e3)

    step_e2e()
    ms_e2e = timed(step_e2e, max(1, min(args.steps, 2)))
    e2e_value = world * B * (n / FRAME_RATE) * max(1, min(args.steps, 2)) / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in pin.values())
    d2h = out_host.numel() * out_host.element_size()

    # ---- profiled pass (rank 0): one Euler update with CUDA events around every launch --------------------------------------
    pk = peaks()
    roof, fwd = None, None
    if rank == 0:
        L = _lib.lib()
        torch.cuda.synchronize()
        L.e2b_prof_enable(1)
        model.sample(torch.zeros_like(res['y0']), text=res['clip'], lens=lens, duration=lens, context=res['ctx'], context_mask=ctx_mask,
                     noise=res['y0'], **dict(kw, steps=2))
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
        cpu_port_seconds_per_update(sd, n, threads)              # warm-up (page-in, thread pool)
        per = cpu_port_seconds_per_update(sd, n, threads, updates=2)
        cpu = dict(value=(n / FRAME_RATE) / (per * (S - 1)), unit=UNIT, cores=threads, kind='port',
                   sample=f'oracle (torch fp32 port of the reference eager sampler), B=1, n={n}: 2 of {S - 1} Euler updates '
                          f'(4 forwards, {per:.2f} s/update) timed and scaled')

    if rank == 0:
        total_flops = eng.flops_per_forward() * (S - 1) * args.steps * world
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16', data='synthetic',
                    config=dict(workload=f'C2: V2A CFM sampling, {B} synthetic {n / FRAME_RATE:.0f} s clips per GPU (CLIP+T5 cond), '
                                         f'{S} grid points ({S - 1} Euler updates) with 2-pass CFG 2.0, shipped 776M 3-stream arch, '
                                         f'random-init weights', clips_per_gpu=B, frames=n, sample_steps=S, guidance_passes=2,
                                parallelism=f'shard-by-clip x{world}' if world > 1 else 'single GPU',
                                l2_policy='working set >> L2 (6 GB of activations streamed per forward); no flush needed'),
                    tflops_executed=total_flops / (ms * 1e-3) / 1e12 / world,
                    tensor_util_vs_sustained=total_flops / (ms * 1e-3) / 1e12 / world / pk['sustained'],
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=launches, clocks=clocks, roofline=roof, forward=fwd, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
