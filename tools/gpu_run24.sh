set -x
mkdir -p gpurun_out
for v in 0 1000000 0 1000000; do
  E2B_OVERLAP_ROWS=$v timeout 600 python bench.py --no-cpu-baseline --steps 2 2> gpurun_out/r2_bench24_$v.err | tee -a gpurun_out/r2_bench24_$v.json | cut -c1-110
done
for v in 0 1000000; do
  E2B_OVERLAP_ROWS=$v timeout 600 python bench.py --no-cpu-baseline --steps 3 --batch 16 2>> gpurun_out/r2_bench24_$v.err | tee -a gpurun_out/r2_bench24_b16_$v.json | cut -c1-110
  E2B_OVERLAP_ROWS=$v timeout 600 python bench.py --no-cpu-baseline --steps 3 --batch 32 2>> gpurun_out/r2_bench24_$v.err | tee -a gpurun_out/r2_bench24_b32_$v.json | cut -c1-110
done
