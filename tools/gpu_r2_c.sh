#!/bin/bash
# round 2, GPU call C (2 GPUs): the driver's scaling launch of the new default bench (independent records + the strong
# sub-record), the 2-GPU parity test, the Lair fib shard mode
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 \
    > gpurun_out/r2c_bench_lurkfib20_2gpu.json 2> gpurun_out/r2c_bench_lurkfib20_2gpu.err
python -m pytest tests -m gpu -x -q -k "two_gpus or device_group" > gpurun_out/r2c_gpu_tests_2gpu.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload fib \
    > gpurun_out/r2c_bench_fib20_2gpu.json 2> gpurun_out/r2c_bench_fib20_2gpu.err
tail -3 gpurun_out/r2c_gpu_tests_2gpu.log; head -c 400 gpurun_out/r2c_bench_lurkfib20_2gpu.json; tail -5 gpurun_out/r2c_bench_lurkfib20_2gpu.err
