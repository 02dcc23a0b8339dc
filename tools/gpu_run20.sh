set -x
mkdir -p gpurun_out
for v in old t64 t32; do
  timeout 300 python tools/ab_lib.py tools/ab/libe2b_$v.so tools/bench_dwconv.py > gpurun_out/r2_dwconv20_$v.txt 2>&1; cat gpurun_out/r2_dwconv20_$v.txt
done
for v in old t64 t32 old t64 t32; do
  timeout 600 python tools/ab_lib.py tools/ab/libe2b_$v.so bench.py --no-cpu-baseline --steps 2 2> gpurun_out/r2_bench20_$v.err | tee -a gpurun_out/r2_bench20_$v.json | cut -c1-120
done
