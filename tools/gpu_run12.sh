set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_1_gemm.py -m gpu -q -x -k "cta_pair" > gpurun_out/r2_tests12.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests12.log
tail -30 gpurun_out/r2_tests12.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_1_gemm.py tests/test_gpu_3_elementwise.py -m gpu -q > gpurun_out/r2_tests12b.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests12b.log
tail -5 gpurun_out/r2_tests12b.log | cut -c1-300
AB=pair timeout 400 python tools/bench_gemm_pf.py > gpurun_out/r2_gemm_pair12.txt 2>&1
cat gpurun_out/r2_gemm_pair12.txt
E2B_GEMM_CG2=1 timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof12_pair.json > gpurun_out/r2_bench12_pair.json 2> gpurun_out/r2_bench12_pair.err
timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof12.json > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err
tail -c 300 gpurun_out/r2_bench12_pair.err
