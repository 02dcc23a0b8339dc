set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_tests25.log 2>&1; tail -6 gpurun_out/r2_tests25.log
