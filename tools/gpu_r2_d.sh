#!/bin/bash
# round 2, GPU call D: full GPU suite (stage entry points, record import), per-launch list of the Lurk fib prove with all
# 41 chips compiled, the other BASELINE configs: lcs at 2^22 rows, fib at 2^24 rows = 4 default-size shards on one GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_lurkfib20.csv \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2d_ncu_launches.log 2>&1
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 > gpurun_out/r2d_bench_lurklcs22.json 2> gpurun_out/r2d_bench_lurklcs22.err
python bench.py --workload lurk-fib --log-rows 22 --shards-per-gpu 4 --steps 2 --no-cpu-baseline > gpurun_out/r2d_bench_lurkfib24.json 2> gpurun_out/r2d_bench_lurkfib24.err
python bench.py --workload lurk-sum --steps 3 --no-cpu-baseline > gpurun_out/r2d_bench_lurksum20.json 2> gpurun_out/r2d_bench_lurksum20.err
tail -3 gpurun_out/r2d_gpu_tests.log; for f in lurklcs22 lurkfib24 lurksum20; do head -c 300 gpurun_out/r2d_bench_$f.json; echo; tail -2 gpurun_out/r2d_bench_$f.err; done
