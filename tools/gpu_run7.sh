set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_1_gemm.py -m gpu -q -x > gpurun_out/r2_tests7.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests7.log
tail -15 gpurun_out/r2_tests7.log
python -m pytest tests/test_gpu_3_elementwise.py tests/test_gpu_4_path.py -m gpu -q > gpurun_out/r2_tests7b.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests7b.log
tail -5 gpurun_out/r2_tests7b.log
timeout 600 python tools/bench_gemm_pf.py > gpurun_out/r2_gemm_resid7.txt 2>&1
cat gpurun_out/r2_gemm_resid7.txt
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/r2_hbm7.txt 2>&1
python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof7.json > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
E2B_RESID_TMA=0 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof7_classic.json > gpurun_out/r2_bench7_classic.json 2> gpurun_out/r2_bench7_classic.err
