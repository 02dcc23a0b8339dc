set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_tests9.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests9.log
tail -12 gpurun_out/r2_tests9.log
python -m pytest tests/test_gpu_9_long.py -m gpu -q -s > gpurun_out/r2_tests9_long.log 2>&1
grep -E "rel-L2|passed|failed" gpurun_out/r2_tests9_long.log
python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof9.json > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
E2B_FUSE_NORM=0 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof9_nofuse.json > gpurun_out/r2_bench9_nofuse.json 2> gpurun_out/r2_bench9_nofuse.err
AB=resid timeout 600 python tools/bench_gemm_pf.py > gpurun_out/r2_gemm_resid9.txt 2>&1
