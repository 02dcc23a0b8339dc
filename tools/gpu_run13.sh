set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_tests13.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests13.log
tail -12 gpurun_out/r2_tests13.log | cut -c1-300
python -m pytest tests/test_gpu_9_long.py -m gpu -q -s > gpurun_out/r2_tests13_long.log 2>&1
grep -E "one sample|C2 batch|passed|failed" gpurun_out/r2_tests13_long.log
AB=pair timeout 400 python tools/bench_gemm_pf.py > gpurun_out/r2_gemm_pair13.txt 2>&1
grep qkv gpurun_out/r2_gemm_pair13.txt
timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof13.json > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err
E2B_GEMM_CG2=0 timeout 600 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof13_nopair.json > gpurun_out/r2_bench13_nopair.json 2> gpurun_out/r2_bench13_nopair.err
