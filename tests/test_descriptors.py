"""The hand-written UMMA shared-memory / instruction descriptor encoders (csrc/ptx.cuh) must agree bit-for-bit with the
CuTe structs shipped in the image (cute/arch/mma_sm100_desc.hpp).  Host-only compile, no GPU needed."""
import glob
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cutlass_include():
    for sp in sys.path:
        for sub in ('flashinfer/data/cutlass/include', 'tilelang/3rdparty/cutlass/include'):
            p = os.path.join(sp, sub)
            if os.path.exists(os.path.join(p, 'cute', 'arch', 'mma_sm100_desc.hpp')):
                return p
    return None


@pytest.mark.skipif(shutil.which('nvcc') is None and not os.path.exists('/usr/local/cuda/bin/nvcc'), reason='nvcc not available')
def test_descriptor_bits_match_cute(tmp_path):
    inc = _cutlass_include()
    if inc is None:
        pytest.skip('CuTe headers not found in this image')
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    exe = str(tmp_path / 'desc_check')
    subprocess.check_call([nvcc, '-std=c++17', '-w', f'-I{inc}', f'-I{ROOT}/video-to-audio-and-piano-rp_b200/csrc',
                           '-gencode', 'arch=compute_100a,code=sm_100a', os.path.join(ROOT, 'tests', 'desc_check.cu'), '-o', exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and 'OK' in out.stdout, out.stdout
