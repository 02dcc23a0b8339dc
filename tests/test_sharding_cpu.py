"""Multi-process (world_size 2, gloo, CPU) test of the shard-by-clip plumbing used for N > 1 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from e2_tts_pytorch.sharding import gather_latents, pad_context_to_common_length, shard_clips
from oracle import synth


def test_shard_clips_partition():
    for n in (1, 7, 64, 512, 513):
        for w in (1, 2, 4, 8):
            parts = [shard_clips(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, num_clips, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mine = shard_clips(num_clips, rank, world)
        # stand-in for the per-rank sampler: a deterministic function of the GLOBAL clip index only
        local = torch.stack([synth.noise(i, 6, 4) * (i + 1) for i in mine]) if len(mine) else torch.zeros(0, 6, 4)
        full = gather_latents(local, num_clips)
        # ragged T5 contexts: rank r holds contexts of r + 3 tokens; every rank must end up with the longest length of all ranks
        ctx, mask = torch.ones(max(len(mine), 1), rank + 3, 8), torch.ones(max(len(mine), 1), rank + 3, dtype=torch.bool)
        ctx, mask = pad_context_to_common_length(ctx, mask)
        assert ctx.shape[1] == mask.shape[1] == world + 2 and bool(mask[:, :rank + 3].all()) and not bool(mask[:, rank + 3:].any())
        assert float(ctx[:, rank + 3:].abs().sum()) == 0.0
        q.put((rank, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('num_clips', [4, 5])
def test_gather_is_world_size_invariant(num_clips):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.stack([synth.noise(i, 6, 4) * (i + 1) for i in range(num_clips)])      # the single-process result
    assert torch.equal(results[0], want) and torch.equal(results[1], want)


def test_pad_context_single_process_is_identity():
    ctx, mask = torch.randn(2, 5, 8), torch.ones(2, 5, dtype=torch.bool)
    c2, m2 = pad_context_to_common_length(ctx, mask)
    assert c2 is ctx and m2 is mask
