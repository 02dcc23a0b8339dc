set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_elementwise.py -m gpu -x -q > gpurun_out/r2_tests18_el.log 2>&1; tail -8 gpurun_out/r2_tests18_el.log
timeout 300 python tools/bench_dwconv.py > gpurun_out/r2_dwconv18.txt 2>&1; cat gpurun_out/r2_dwconv18.txt
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_1_gemm.py --deselect tests/test_gpu_3_elementwise.py > gpurun_out/r2_tests18.log 2>&1; tail -5 gpurun_out/r2_tests18.log
timeout 600 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof18.json > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err
cut -c1-200 gpurun_out/r2_bench18.json
