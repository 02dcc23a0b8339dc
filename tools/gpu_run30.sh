mkdir -p gpurun_out
for v in full nofma nosilu neither; do
  echo "== $v"; timeout 120 python tools/ab_lib.py tools/ab/libe2b_$v.so tools/bench_dwconv.py 2>&1 | grep -v "rel err" | head -3
done | tee gpurun_out/r2_dw30.txt
