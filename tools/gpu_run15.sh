set -x
mkdir -p gpurun_out
E2B_GEMM_CG2_SKIP=1024x1024 timeout 600 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof15_skip.json > gpurun_out/r2_bench15_skip.json 2> gpurun_out/r2_bench15_skip.err
timeout 600 python bench.py --no-cpu-baseline --steps 2 > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err
python tools/prof_forward.py --batch 64 > gpurun_out/r2_ncu15_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"e2b" -c 900 --csv \
    --log-file gpurun_out/r2_ncu15_launches.csv python tools/prof_forward.py --batch 64 > gpurun_out/r2_ncu15.log 2>&1
python tools/ncu_tensor_share.py gpurun_out/r2_ncu15_launches.csv > gpurun_out/r2_ncu15_tensor_share.txt 2>&1
cat gpurun_out/r2_ncu15_tensor_share.txt
ls -la gpurun_out/r2_ncu15*
