set -x
mkdir -p gpurun_out
timeout 300 python bench.py --batch 1 --no-cpu-baseline --steps 5 --profile-out gpurun_out/r2_prof22_b1.json > gpurun_out/r2_bench22_b1.json 2> gpurun_out/r2_bench22_b1.err
timeout 300 python bench.py --batch 4 --no-cpu-baseline --steps 5 --profile-out gpurun_out/r2_prof22_b4.json > gpurun_out/r2_bench22_b4.json 2> gpurun_out/r2_bench22_b4.err
cut -c1-300 gpurun_out/r2_bench22_b1.json
