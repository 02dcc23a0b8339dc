set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_elementwise.py tests/test_gpu_4_path.py -m gpu -x -q > gpurun_out/r2_tests19_el.log 2>&1; tail -4 gpurun_out/r2_tests19_el.log
timeout 300 python tools/bench_dwconv.py > gpurun_out/r2_dwconv19.txt 2>&1; cat gpurun_out/r2_dwconv19.txt
timeout 600 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof19.json > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err
cut -c1-200 gpurun_out/r2_bench19.json
