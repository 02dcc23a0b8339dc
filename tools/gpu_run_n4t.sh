set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_12_multigpu.py -m gpu -q -s > gpurun_out/r02_gpu_tests_4gpu_shard_equivalence.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_gpu_tests_4gpu_shard_equivalence.log
tail -6 gpurun_out/r02_gpu_tests_4gpu_shard_equivalence.log
