"""Stage the reference files the checker imports into the git-ignored baseline/_ref/ so that they travel to the GPU box
(gpurun ships the working tree minus .git; /root/reference does not exist there).  Run in the build container:

    python -m oracle.stage_reference

Copies, byte for byte and outside of git history (baseline/_ref/ is listed in .gitignore, not in .gpurunignore):
  src/e2_tts_pytorch/e2_tts_crossatt3.py     the reference's model / sampler module ("X3")
  src/e2_tts_pytorch/__init__.py             (if present)
  src/audeo/Video2RollNet.py                 imported by X3 at module level (X3:55-57)
With them present `oracle.ref_loader.reference_available()` is true on the box: `bench.py --impl reference` and the
`cpu_baseline` leg time the reference's OWN E2TTS.sample (kind "reference") instead of the oracle port, and the
oracle-vs-reference tests run there too.  Nothing under baseline/_ref is imported by the product path."""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')
FILES = ['src/e2_tts_pytorch/e2_tts_crossatt3.py', 'src/e2_tts_pytorch/__init__.py', 'src/audeo/Video2RollNet.py']


def stage(verbose=True):
    if not os.path.isdir(SRC):
        return False
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.exists(s):
            if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s) or os.path.getsize(d) != os.path.getsize(s):
                shutil.copyfile(s, d)
        elif rel.endswith('__init__.py') and not os.path.exists(d):
            open(d, 'w').close()
    if verbose:
        print('staged', len(FILES), 'reference files under', DST)
    return True


if __name__ == '__main__':
    stage()
