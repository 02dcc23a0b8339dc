set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_tests10.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests10.log
tail -12 gpurun_out/r2_tests10.log
python -m pytest tests/test_gpu_9_long.py -m gpu -q -s > gpurun_out/r2_tests10_long.log 2>&1
grep -E "rel-L2|passed|failed" gpurun_out/r2_tests10_long.log
python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof10.json > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err
E2B_FUSE_NORM=0 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof10_nofuse.json > gpurun_out/r2_bench10_nofuse.json 2> gpurun_out/r2_bench10_nofuse.err
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/r2_hbm10.txt 2>&1
timeout 300 python tools/bench_dwconv.py > gpurun_out/r2_dwconv10.txt 2>&1
