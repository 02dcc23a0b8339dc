set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_1_gemm.py tests/test_gpu_2_attention.py tests/test_gpu_4_path.py -m gpu -q > gpurun_out/r2_tests5.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests5.log
tail -3 gpurun_out/r2_tests5.log
timeout 600 python tools/bench_attention.py > gpurun_out/r2_attn_ab5.txt 2>&1
timeout 600 python tools/bench_gemm_pf.py > gpurun_out/r2_gemm_pf5.txt 2>&1
python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof5.json > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
E2B_GEMM_PF=0 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof5_nopf.json > gpurun_out/r2_bench5_nopf.json 2> gpurun_out/r2_bench5_nopf.err
python tools/prof_forward.py --batch 64 > gpurun_out/r2_ncu5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attention_kernel|dwconv|rmsnorm" -s 14 -c 31 -o gpurun_out/r2_prof_layer python tools/prof_forward.py --batch 64 > gpurun_out/r2_ncu5.log 2>&1
tail -3 gpurun_out/r2_ncu5.log
ls -la gpurun_out
