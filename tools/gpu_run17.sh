set -x
mkdir -p gpurun_out
python tools/prof_forward.py --batch 64 > gpurun_out/r2_ncu17_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attention_kernel|dwconv_tma_kernel" -s 14 -c 24 -o /tmp/r02_layer python tools/prof_forward.py --batch 64 > gpurun_out/r2_ncu17_full.log 2>&1
python tools/ncu_summary.py /tmp/r02_layer.ncu-rep 10 > gpurun_out/r2_ncu17_layer_summary.txt 2>&1
ls -la /tmp/r02_layer.ncu-rep; head -c 3000 gpurun_out/r2_ncu17_layer_summary.txt
