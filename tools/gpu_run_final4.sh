set -x
mkdir -p gpurun_out
timeout 900 python bench.py --profile-out gpurun_out/r02_event_profile.json > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --config C4 --no-cpu-baseline > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
timeout 600 python bench.py --config C5 --no-cpu-baseline --steps 2 > gpurun_out/r02_bench_c5_s64_n1500.json 2> gpurun_out/r02_bench_c5.err
timeout 600 python bench.py --config C5 --no-cpu-baseline --sample-steps 16 --frames 375 --batch 64 > gpurun_out/r02_bench_c5_s16_n375.json 2>> gpurun_out/r02_bench_c5.err
cut -c1-160 gpurun_out/r02_bench_n1.json gpurun_out/r02_bench_c4.json gpurun_out/r02_bench_c5_s64_n1500.json gpurun_out/r02_bench_c5_s16_n375.json
