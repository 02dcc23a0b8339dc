set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_1_gemm.py -m gpu -x -q > gpurun_out/r2_tests16_gemm.log 2>&1; tail -15 gpurun_out/r2_tests16_gemm.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_1_gemm.py > gpurun_out/r2_tests16.log 2>&1; tail -5 gpurun_out/r2_tests16.log
AB=pair timeout 600 python tools/bench_gemm_pf.py 2>&1 | grep qkv > gpurun_out/r2_gemm_qkv16.txt
E2B_QKV_TMA=0 AB=pair timeout 600 python tools/bench_gemm_pf.py 2>&1 | grep qkv > gpurun_out/r2_gemm_qkv16_classic.txt
cat gpurun_out/r2_gemm_qkv16.txt gpurun_out/r2_gemm_qkv16_classic.txt
timeout 600 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof16.json > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err
E2B_QKV_TMA=0 timeout 600 python bench.py --no-cpu-baseline --steps 2 > gpurun_out/r2_bench16_classic.json 2> gpurun_out/r2_bench16_classic.err
cat gpurun_out/r2_bench16.json gpurun_out/r2_bench16_classic.json | cut -c1-200
