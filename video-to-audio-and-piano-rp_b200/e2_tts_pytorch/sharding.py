"""Shard-by-clip plumbing for multi-GPU sampling (one process per GPU, torch.distributed).

Each clip's ODE is independent, so ranks never exchange anything inside the sampling loop; the only
collective is one gather of the finished latents.  Per-clip synthetic seeds / conditions are keyed by the GLOBAL clip index,
which makes results invariant to the world size.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_clips(num_clips: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced block of global clip indices owned by `rank` (first `num_clips % world_size` ranks get one more)."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of size {world_size}')
    base, extra = divmod(num_clips, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_latents(local: torch.Tensor, num_clips: int, group=None) -> torch.Tensor:
    """All-gather `[B_local, n, d]` latents into `[num_clips, n, d]` in global clip order (ragged shards are padded)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [len(shard_clips(num_clips, r, world)) for r in range(world)]
    bmax = max(sizes)
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    padded = local if local.shape[0] == bmax else torch.cat([local, local.new_zeros(bmax - local.shape[0], *local.shape[1:])])
    out = local.new_empty(world * bmax, *local.shape[1:])
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return torch.cat([out[r * bmax:r * bmax + sizes[r]] for r in range(world)])


def pad_context_to_common_length(context: torch.Tensor, context_mask: torch.Tensor, group=None):
    """Pad a rank's T5 context batch `[B_local, nc_local, d]` (+ mask `[B_local, nc_local]`) to the longest context of ALL ranks.

    The reference rotates cross-attention keys with the LAST `nc` rows of the rotary table (x-transformers
    `apply_rotary_pos_emb`: `freqs[:, -seq_len:]`, restated in oracle/third_party.py), so a clip's latent depends on the padded
    length of the context batch it is sampled with -- in the reference as in this drop-in.  A sharded run therefore equals the
    one-process run per clip only if every shard pads to the same length; this helper makes that so with one MAX all-reduce.
    """
    nc = torch.tensor([context.shape[1]], dtype=torch.int64, device=context.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(nc, op=dist.ReduceOp.MAX, group=group)
    pad = int(nc.item()) - context.shape[1]
    if pad > 0:
        context = torch.cat([context, context.new_zeros(context.shape[0], pad, context.shape[2])], dim=1)
        context_mask = torch.cat([context_mask, context_mask.new_zeros(context_mask.shape[0], pad)], dim=1)
    return context, context_mask
