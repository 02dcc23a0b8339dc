set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_2_attention.py tests/test_gpu_4_path.py -m gpu -q -s > gpurun_out/r2_tests2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests2.log
timeout 600 python tools/bench_attention.py > gpurun_out/r2_attn_ab.txt 2>&1
tail -12 gpurun_out/r2_attn_ab.txt
python bench.py --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -3 gpurun_out/r2_tests2.log
