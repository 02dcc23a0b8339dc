# Second part of the round-2 evidence run: ncu launch list of one Euler update (weight packing kernels filtered out) and ncu --set full of the
# sampler epilogue / roll front-end kernels
set -x
mkdir -p gpurun_out
python tools/prof_forward.py --batch 64 > gpurun_out/r02_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
    -k regex:"gemm_kernel|attention_kernel|dwconv_tma_kernel|rmsnorm_kernel|guided_euler_kernel|init_stream_kernel|cast_pad_kernel|time_mlp_kernel|time_gemv_kernel|apg_reduce_kernel|mask_rows_kernel" \
    -c 900 --csv --log-file gpurun_out/r02_ncu_launches.csv python tools/prof_forward.py --batch 64 > gpurun_out/r02_ncu_launches.log 2>&1
python tools/ncu_tensor_share.py gpurun_out/r02_ncu_launches.csv > gpurun_out/r02_ncu_tensor_share.txt 2>&1
cat gpurun_out/r02_ncu_tensor_share.txt
ncu --set full --clock-control none --import-source on -k regex:"guided_euler_kernel|frame_windows_kernel|roll_expand_kernel|stage_clip_kernel|apg_reduce_kernel" -c 8 -o /tmp/r02_hbm2 python tools/bench_hbm_kernels.py > gpurun_out/r02_ncu_hbm2.log 2>&1
python tools/ncu_summary.py /tmp/r02_hbm2.ncu-rep 6 > gpurun_out/r02_ncu_hbm_kernels_summary2.txt 2>&1
cut -c1-300 gpurun_out/r02_ncu_hbm_kernels_summary2.txt | head -30
