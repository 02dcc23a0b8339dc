set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_gpu_tests_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_gpu_tests_final.log
tail -4 gpurun_out/r02_gpu_tests_final.log
