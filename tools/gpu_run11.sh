set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_1_gemm.py tests/test_gpu_3_elementwise.py tests/test_gpu_4_path.py -m gpu -q > gpurun_out/r2_tests11.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests11.log
tail -5 gpurun_out/r2_tests11.log
timeout 300 python tools/bench_dwconv.py > gpurun_out/r2_dwconv11.txt 2>&1
cat gpurun_out/r2_dwconv11.txt
python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof11.json > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
E2B_FUSE_CONV=0 python bench.py --no-cpu-baseline --profile-out gpurun_out/r2_prof11_noconv.json > gpurun_out/r2_bench11_noconv.json 2> gpurun_out/r2_bench11_noconv.err
E2B_FUSE_CONV=0 E2B_RT_EW8_MAXK=2304 python bench.py --no-cpu-baseline --steps 2 --profile-out gpurun_out/r2_prof11_k2304.json > gpurun_out/r2_bench11_k2304.json 2> gpurun_out/r2_bench11_k2304.err
