set -x
mkdir -p gpurun_out
AB=pair timeout 600 python tools/bench_gemm_pf.py > gpurun_out/r2_gemm_pair14.txt 2>&1
tail -8 gpurun_out/r2_gemm_pair14.txt
