set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_1_gemm.py tests/test_gpu_2_attention.py tests/test_gpu_3_elementwise.py tests/test_gpu_4_path.py -m gpu -q > gpurun_out/r2_tests6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests6.log
tail -3 gpurun_out/r2_tests6.log
timeout 600 python tools/bench_attention.py > gpurun_out/r2_attn_ab6.txt 2>&1
timeout 600 python tools/bench_gemm_pf.py > gpurun_out/r2_gemm_pf6.txt 2>&1
timeout 600 python tools/bench_hbm_kernels.py > gpurun_out/r2_hbm6.txt 2>&1
python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof6.json > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
E2B_GEMM_PF=0 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof6_nopf.json > gpurun_out/r2_bench6_nopf.json 2> gpurun_out/r2_bench6_nopf.err
python tools/prof_forward.py --batch 64 > gpurun_out/r2_ncu6_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attention_kernel|dwconv|rmsnorm" -s 14 -c 31 -o /tmp/r2_prof_layer python tools/prof_forward.py --batch 64 > gpurun_out/r2_ncu6.log 2>&1
python tools/ncu_summary.py /tmp/r2_prof_layer.ncu-rep 8 > gpurun_out/r2_ncu6_summary.txt 2>&1
ncu -i /tmp/r2_prof_layer.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null; python - <<'PY' > gpurun_out/r2_ncu6_raw_small.csv
import csv
rows = list(csv.reader(open('/tmp/raw.csv')))
keep = ['ID', 'Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']
idx = [rows[0].index(k) for k in keep if k in rows[0]]
w = csv.writer(__import__('sys').stdout)
for r in rows:
    w.writerow([r[i][:70] for i in idx])
PY
ls -la gpurun_out /tmp/r2_prof_layer.ncu-rep
