# 4 GPUs of one box: shard-equivalence test at 2 ranks (needs >= 2 GPUs), then the default bench at N = 4 and N = 2 on the same box
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n4_gpus.txt
python -m pytest tests/test_gpu_12_multigpu.py -m gpu -q -s > gpurun_out/r02_gpu_tests_2gpu_shard_equivalence.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_gpu_tests_2gpu_shard_equivalence.log
tail -4 gpurun_out/r02_gpu_tests_2gpu_shard_equivalence.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
tail -c 700 gpurun_out/r02_bench_n4.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 700 gpurun_out/r02_bench_n2.json
