set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt
python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2_tests1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests1.log
python -m pytest tests/test_gpu_9_long.py tests/test_gpu_10_inpaint.py -m gpu -q -s > gpurun_out/r2_tests1_long.log 2>&1
python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
python bench.py --config C4 --no-cpu-baseline > gpurun_out/r2_bench1_c4.json 2> gpurun_out/r2_bench1_c4.err
python bench.py --config C5 --no-cpu-baseline --steps 2 > gpurun_out/r2_bench1_c5.json 2> gpurun_out/r2_bench1_c5.err
tail -5 gpurun_out/r2_tests1.log
