"""bench.py host logic that needs no GPU: the BASELINE.json configurations it offers, the committed ncu traffic figures it
reports as roofline.traffic, the measured-peak lookup, the flag surface of the driver's contract, and the loud failure of the
GPU arm on a box without CUDA (there is no CPU path behind it)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_configs_are_the_baseline_ones():
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert bench.METRIC.startswith('generated audio-sec/sec') and base['metric'].startswith('generated audio-sec/sec')
    assert sorted(bench.CONFIGS) == ['C2', 'C3', 'C4', 'C5']
    c2, c3, c4, c5 = (bench.CONFIGS[k] for k in ('C2', 'C3', 'C4', 'C5'))
    assert (c2['batch'], c2['frames'], c2['sample_steps'], c2['guidance']) == (64, 750, 32, None)       # 64 x 10 s, 32 steps, CFG
    assert c3['total'] == 512 and c3['frames'] == 750                                                   # 512 clips sharded by clip
    assert c4['frames'] == 2250 and c4['live_roll']                                                     # 30 s piano clips, roll live
    assert c5['guidance'] and len(c5['guidance']) == 3 and c5['sample_steps'] == 64                     # K-pass guidance
    for c in bench.CONFIGS.values():
        assert '{B}' in c['name'] and c['frames'] / bench.FRAME_RATE in (10.0, 20.0, 30.0)


def test_committed_ncu_traffic_is_loadable_and_matches_the_kernel_keys():
    tr = bench._load_traffic()
    assert tr, 'profiles/r0*_ncu_traffic.json missing'
    for key, v in tr.items():
        kind, m, n, k = key.split()
        assert kind in ('attention', 'gemm_geglu', 'gemm_resid', 'gemm_qkv') and m.startswith('M=') and n.startswith('N=') and k.startswith('K=')
        assert v['dram_bytes_per_launch'] > 1e6 and v['source'].startswith('profiles/')
        assert os.path.exists(os.path.join(ROOT, v['source'].split()[0].rstrip(':')))


def test_peaks_come_from_the_driver_file_or_the_stated_fallback():
    pk = bench.peaks()
    assert pk['burst'] >= pk['sustained'] > 100 and pk['src']
    if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
        m = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        assert pk['burst'] == pytest.approx(m['bf16_tflops']) and 'measured' in pk['src']


def test_flag_surface_of_the_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--help'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0
    for flag in ('--gpus', '--steps', '--warmup', '--impl', '--config'):
        assert flag in out.stdout


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_gpu_arm_fails_loudly_without_cuda():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode != 0 and 'CUDA' in out.stderr and not out.stdout.strip().startswith('{')
