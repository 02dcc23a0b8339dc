# Round-2 evidence, final build (branch overlap on): bench lines, event profile, small batches, HBM kernels
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --profile-out gpurun_out/r02_event_profile.json > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --config C4 --no-cpu-baseline > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
timeout 600 python bench.py --config C5 --no-cpu-baseline --steps 2 > gpurun_out/r02_bench_c5_s64_n1500.json 2> gpurun_out/r02_bench_c5.err
timeout 600 python bench.py --config C5 --no-cpu-baseline --sample-steps 16 --frames 375 --batch 64 > gpurun_out/r02_bench_c5_s16_n375.json 2>> gpurun_out/r02_bench_c5.err
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/r02_hbm_kernels.txt 2>&1
timeout 300 python tools/bench_latency.py > gpurun_out/r02_latency.txt 2>&1
python -m pytest tests/test_gpu_9_long.py tests/test_gpu_10_inpaint.py -m gpu -q -s > gpurun_out/r02_parity_long_trajectories.log 2>&1
tail -3 gpurun_out/r02_parity_long_trajectories.log
cut -c1-160 gpurun_out/r02_bench_n1.json gpurun_out/r02_bench_c4.json gpurun_out/r02_bench_c5_s64_n1500.json gpurun_out/r02_bench_c5_s16_n375.json
cat gpurun_out/r02_latency.txt gpurun_out/r02_hbm_kernels.txt
