set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_1_gemm.py -m gpu -q > gpurun_out/r2_tests8.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests8.log
tail -15 gpurun_out/r2_tests8.log
python -m pytest tests/test_gpu_9_long.py tests/test_gpu_5_configs.py tests/test_gpu_4_path.py -m gpu -q -s > gpurun_out/r2_tests8b.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests8b.log
grep -E "rel-L2|passed|failed" gpurun_out/r2_tests8b.log | tail -20
