#!/bin/bash
# round 2, GPU call L (8 GPUs): the driver's scaling launch at N=8 (independent records + strong sub-record: one Lurk record
# proved by all eight GPUs), column-sharded prove of one 2^24-row Lair fib shard (round 1: 64.8 ms), Lurk fib 2^24 as four
# default-size shards over 4 pairs of GPUs
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_lurkfib20_8gpu.json 2> gpurun_out/r2l_bench_lurkfib20_8gpu.err
SHARD_SIZE=16777216 timeout 600 $T --master-port 29512 tools/column_sharded_prove.py --log-rows 24 --skip-parity-set --steps 3 > gpurun_out/r2l_column_fib24_8gpu.json 2> gpurun_out/r2l_column_fib24_8gpu.err
timeout 900 $T --master-port 29513 bench.py --gpus 8 --workload lurk-fib --log-rows 22 --partition column --gpus-per-shard 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_lurkfib24_8gpu_4x2.json 2> gpurun_out/r2l_bench_lurkfib24_8gpu_4x2.err
timeout 600 $T --master-port 29514 bench.py --gpus 8 --workload fib --log-rows 22 --partition column --gpus-per-shard 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_fib24_8gpu_4x2.json 2> gpurun_out/r2l_bench_fib24_8gpu_4x2.err
for f in bench_lurkfib20_8gpu column_fib24_8gpu bench_lurkfib24_8gpu_4x2 bench_fib24_8gpu_4x2; do echo "== $f"; head -c 400 gpurun_out/r2l_$f.json; echo; tail -2 gpurun_out/r2l_$f.err | cut -c1-300; done
