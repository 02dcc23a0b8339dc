"""Multi-GPU shard equivalence on real devices (SURVEY.md section 4 iv): an N-rank run (shard by clip + one NCCL gather) must
equal the 1-rank run per clip, bit for bit.  Needs >= 2 GPUs on the box (`gpurun --gpus 2 -- python -m pytest tests -m gpu -k
multigpu`); skips on a single-GPU box (where tests/test_sharding_cpu.py covers the bookkeeping with gloo)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _launch(world, arch, clips, n, steps, port):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(HERE, 'multigpu_worker.py'), arch, str(clips), str(n), str(steps)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-12000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize('arch,clips,n,steps', [('tiny', 9, 60, 5), ('shipped', 8, 750, 3)])
def test_n_rank_equals_one_rank_per_clip(arch, clips, n, steps):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if ngpu < 4 else 4
    r = _launch(world, arch, clips, n, steps, 29650 + (0 if arch == 'tiny' else 1))
    print(r)
    assert r['ok'] and r['world'] == world, r
