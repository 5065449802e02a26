#!/bin/bash
# round 2, GPU call I (2 GPUs): K2 on stage-native twiddles (no inter-pass products), device-group record upload through one
# arena + a single all-gather
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "lde or coset or merkle or lurk or two_gpus or device_group or bit_exact or stage or sharded" > gpurun_out/r2i_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2i_bench_lurkfib20.json 2> gpurun_out/r2i_bench_lurkfib20.err
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2i_bench_lurklcs22.json 2> gpurun_out/r2i_bench_lurklcs22.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r2i_bench_lurkfib20_2gpu.json 2> gpurun_out/r2i_bench_lurkfib20_2gpu.err
ncu --set full --clock-control none -k regex:lde_k2 --launch-skip 3 -c 2 -o gpurun_out/r2i_k2 \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2i_ncu_k2.log 2>&1
tail -3 gpurun_out/r2i_gpu_tests.log; head -c 300 gpurun_out/r2i_bench_lurkfib20.json; echo; head -c 300 gpurun_out/r2i_bench_lurklcs22.json; echo;  head -c 300 gpurun_out/r2i_bench_lurkfib20_2gpu.json; tail -3 gpurun_out/r2i_bench_lurkfib20_2gpu.err
