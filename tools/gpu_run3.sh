set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_1_gemm.py tests/test_gpu_2_attention.py tests/test_gpu_4_path.py tests/test_gpu_5_configs.py tests/test_gpu_9_long.py -m gpu -q -s > gpurun_out/r2_tests3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests3.log
timeout 600 python tools/bench_attention.py > gpurun_out/r2_attn_ab3.txt 2>&1
tail -12 gpurun_out/r2_attn_ab3.txt
python bench.py --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
E2B_VT=1 python bench.py --no-cpu-baseline --steps 2 > gpurun_out/r2_bench3_vt.json 2> gpurun_out/r2_bench3_vt.err
grep -E "passed|failed|error" gpurun_out/r2_tests3.log | tail -3
