set -x
mkdir -p gpurun_out
python tools/ab_lib.py tools/ab/libe2b_packed64.so tools/bench_dwconv.py > gpurun_out/r2_dw28_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dwconv_tma_kernel" -s 2 -c 2 -o /tmp/dw28 python tools/ab_lib.py tools/ab/libe2b_packed64.so tools/bench_dwconv.py > gpurun_out/r2_dw28_ncu.log 2>&1
python tools/ncu_summary.py /tmp/dw28.ncu-rep 25 > gpurun_out/r2_dw28_summary.txt 2>&1
cut -c1-250 gpurun_out/r2_dw28_summary.txt | head -50
