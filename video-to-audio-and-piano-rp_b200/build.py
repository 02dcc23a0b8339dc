"""Build libe2b.so (sm_100a only) in-tree with nvcc.  `python build.py [--force]`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libe2b.so')
SOURCES = ['gemm.cu', 'attention.cu', 'attention_v1.cu', 'attention_f32.cu', 'elementwise.cu', 'melspec.cu', 'staging.cu', 'frames.cu', 'encodec.cu', 'engine.cu', 'prof.cu']
HEADERS = ['ptx.cuh', 'kernels.h', 'prof.h', os.path.join('..', '..', 'include', 'e2b.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, 'build', s.replace('.cu', '.o'))
        cmd = [_nvcc(), *NVCC_FLAGS, '-c', os.path.join(CSRC, s), '-o', o]
        if verbose:
            cmd.insert(1, '-Xptxas'); cmd.insert(2, '-v')
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f'nvcc failed on {s}')
    cmd = [_nvcc(), '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static', '-ldl']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
