set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_4_path.py -m gpu -x -q > gpurun_out/r2_tests23.log 2>&1; tail -15 gpurun_out/r2_tests23.log
timeout 300 python tools/bench_latency.py > gpurun_out/r2_latency23_overlap.txt 2>&1; cat gpurun_out/r2_latency23_overlap.txt
E2B_OVERLAP_ROWS=0 timeout 300 python tools/bench_latency.py > gpurun_out/r2_latency23_one_stream.txt 2>&1; cat gpurun_out/r2_latency23_one_stream.txt
