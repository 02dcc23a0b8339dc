set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_n2_gpus.txt
python -m pytest tests/test_gpu_12_multigpu.py -m gpu -q -s > gpurun_out/r2_tests_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_tests_n2.log
tail -8 gpurun_out/r2_tests_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 600 gpurun_out/r2_bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --config C3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_c3_n2.json 2> gpurun_out/r2_bench_c3_n2.err
tail -c 400 gpurun_out/r2_bench_c3_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err
tail -c 300 gpurun_out/r2_bench_ref_n2.json
