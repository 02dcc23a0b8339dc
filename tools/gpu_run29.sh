set -x
mkdir -p gpurun_out
timeout 300 python tools/ab_lib.py tools/ab/libe2b_desync64.so tools/pytest_main.py tests/test_gpu_3_elementwise.py -m gpu -q -x -k dwconv > gpurun_out/r2_dw29_tests64.log 2>&1; tail -3 gpurun_out/r2_dw29_tests64.log
timeout 300 python tools/ab_lib.py tools/ab/libe2b_desync32.so tools/pytest_main.py tests/test_gpu_3_elementwise.py -m gpu -q -x -k dwconv > gpurun_out/r2_dw29_tests32.log 2>&1; tail -3 gpurun_out/r2_dw29_tests32.log
for v in base desync64 desync32 base desync64; do
  timeout 120 python tools/ab_lib.py tools/ab/libe2b_$v.so tools/bench_dwconv.py > gpurun_out/r2_dw29_$v.txt 2>&1; echo $v; cat gpurun_out/r2_dw29_$v.txt
done
