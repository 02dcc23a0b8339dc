# Round-2 evidence run (one B200): every number under profiles/r02_* comes from this command list.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_gpu_tests_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_gpu_tests_final.log
tail -5 gpurun_out/r02_gpu_tests_final.log
python -m pytest tests/test_gpu_9_long.py tests/test_gpu_10_inpaint.py -m gpu -q -s > gpurun_out/r02_parity_long_trajectories.log 2>&1
timeout 900 python bench.py --profile-out gpurun_out/r02_event_profile.json > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 600 python bench.py --config C4 --no-cpu-baseline > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
timeout 600 python bench.py --config C5 --no-cpu-baseline --steps 2 > gpurun_out/r02_bench_c5_s64_n1500.json 2> gpurun_out/r02_bench_c5.err
timeout 600 python bench.py --config C5 --no-cpu-baseline --sample-steps 16 --frames 375 --batch 64 > gpurun_out/r02_bench_c5_s16_n375.json 2>> gpurun_out/r02_bench_c5.err
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/r02_hbm_kernels.txt 2>&1
timeout 300 python tools/bench_latency.py > gpurun_out/r02_latency.txt 2>&1
timeout 300 python tools/bench_attention.py > gpurun_out/r02_attention_ab_final.txt 2>&1
# ncu: launch list of one Euler update with the tensor-pipe metric (cold-cache, serialised: shares, not absolutes)
python tools/prof_forward.py --batch 64 > gpurun_out/r02_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 900 --csv \
    --log-file gpurun_out/r02_ncu_launches.csv python tools/prof_forward.py --batch 64 > gpurun_out/r02_ncu_launches.log 2>&1
python tools/ncu_tensor_share.py gpurun_out/r02_ncu_launches.csv > gpurun_out/r02_ncu_tensor_share.txt 2>&1
cat gpurun_out/r02_ncu_tensor_share.txt
# ncu --set full on the HBM-bound front-end / sampler kernels (melspec, guided Euler update, frame windows, condition staging)
ncu --set full --clock-control none --import-source on -k regex:"melspec_kernel|mel_ranges_kernel|guided_euler_kernel|frame_windows_kernel|roll_expand_kernel" -c 10 -o /tmp/r02_hbm python tools/bench_hbm_kernels.py > gpurun_out/r02_ncu_hbm.log 2>&1
python tools/ncu_summary.py /tmp/r02_hbm.ncu-rep 6 > gpurun_out/r02_ncu_hbm_kernels_summary.txt 2>&1
ls -la gpurun_out | tail -30
