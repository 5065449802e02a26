#!/bin/bash
# round 2, GPU call G (4 GPUs): the 2-GPU column-sharded parity test, the driver's scaling launch at N=4 (independent
# records + the strong sub-record: one record proved by all four GPUs), Poseidon2 sweep 2^16..2^28 at N=4 (configs[4])
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "two_gpus or device_group" > gpurun_out/r2g_gpu_tests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_gpu_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r2g_bench_lurkfib20_4gpu.json 2> gpurun_out/r2g_bench_lurkfib20_4gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tools/poseidon2_sweep.py --min-log 16 --max-log 28 \
    --out gpurun_out/r2g_poseidon2_sweep_4gpu.jsonl > gpurun_out/r2g_sweep4.log 2>&1
tail -3 gpurun_out/r2g_gpu_tests_2gpu.log; head -c 300 gpurun_out/r2g_bench_lurkfib20_4gpu.json; echo; tail -3 gpurun_out/r2g_bench_lurkfib20_4gpu.err; tail -3 gpurun_out/r2g_sweep4.log | cut -c1-300
