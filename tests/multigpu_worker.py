"""Worker for tests/test_gpu_12_multigpu.py (launched with torchrun, one process per GPU): every rank samples its shard of a
global clip list (sharding.shard_clips, per-clip inputs keyed by global clip index), the latents are gathered with NCCL
(sharding.gather_latents), and rank 0 compares the gathered result bit for bit with its own single-process run over ALL clips
(SURVEY.md section 4 iv: N-GPU result == 1-GPU result per clip)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import synthetic as synth
from e2_tts_pytorch.e2_tts_crossatt3 import E2TTS
from e2_tts_pytorch.sharding import gather_latents, shard_clips


def main():
    arch, num_clips, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    cfg = synth.SHIPPED if arch == 'shipped' else synth.TINY
    tr = dict(depth=cfg['depth'], dim=cfg['dim'], dim_text=cfg['dim_text'], dim_frames=cfg['dim_frames'], heads=cfg['heads'], dim_head=64,
              max_seq_len=cfg['max_seq_len'], if_text_modules=True, if_cross_attn=True, if_audio_conv=True, if_text_conv=True)
    m = E2TTS(duration_predictor=None, transformer=tr, if_cond_proj_in=False, if_embed_text=False, if_text_encoder2=False, if_clip_encoder=False,
              num_channels=cfg['num_channels'], sampling_rate=24000)
    m.load_state_dict(synth.random_state_dict(**cfg, seed=0), strict=True)
    m = m.to(dev)

    def run(clips):
        # ragged lengths keyed by the GLOBAL clip index only; even clips keep the full length, and every shard of the
        # contiguous partition used by the test holds an even clip, so every batch is padded to the same n
        lens = [n if c % 2 == 0 else n - 5 * (c % 3 + 1) for c in clips]
        assert max(lens) == n
        # Context lengths: ragged too, but every batch must be padded to the SAME context length (even clips keep 8 tokens): the
        # reference rotates cross-attention keys with the LAST rows of the rotary table (x-transformers, X3 via oracle/third_party.py),
        # so a clip's result depends on the padded length of the T5 batch it sits in -- in the reference as in this drop-in.
        bt = synth.batch(clips, n, lens=lens, nc_list=[8 if c % 2 == 0 else 6 + c % 3 for c in clips], dim_text=cfg['dim_text'], dim=cfg['dim'],
                         d=cfg['num_channels'], live_frames=True)
        d = {k: v.to(dev) for k, v in bt.items()}
        out = m.sample(torch.zeros_like(d['y0']), text=d['clip'], lens=d['lens'], duration=d['lens'], steps=steps, cfg_strength=2.0,
                       remove_parallel_component=False, return_raw_output=True, context=d['ctx'], context_mask=d['ctx_mask'],
                       frames=d['frames'], noise=d['y0'])
        return out, lens

    mine = list(shard_clips(num_clips, rank, world))
    local_out, _ = run(mine)
    gathered = gather_latents(local_out, num_clips)
    ok, detail = True, ''
    if rank == 0:
        full, lens = run(list(range(num_clips)))
        assert gathered.shape == full.shape, (gathered.shape, full.shape)
        for c in range(num_clips):                                    # valid rows of every clip, bit for bit
            if not torch.equal(gathered[c, :lens[c]], full[c, :lens[c]]):
                ok = False
                detail += f' clip {c} differs (max abs {float((gathered[c, :lens[c]] - full[c, :lens[c]]).abs().max()):.3e});'
        print(json.dumps(dict(ok=ok, world=world, clips=num_clips, arch=arch, n=n, steps=steps, detail=detail)), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
