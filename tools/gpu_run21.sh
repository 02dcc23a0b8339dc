set -x
mkdir -p gpurun_out
for v in base rowmajor base rowmajor; do
  timeout 300 python tools/ab_lib.py tools/ab/libe2b_$v.so tools/bench_dwconv.py > gpurun_out/r2_dwconv21_$v.txt 2>&1; cat gpurun_out/r2_dwconv21_$v.txt
done
