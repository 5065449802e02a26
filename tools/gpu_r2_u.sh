#!/bin/bash
# round 2, GPU call U: the evidence set of the final code (reduced openings at 4 columns in flight, opened values per round and height) (K2 per size at 128 registers, one-row compiled quotient,
# tiny chips on the side stream, transcript behind the openings): full GPU suite, default bench with CPU leg, launch list
# of the bench command, DRAM bytes of the Merkle launches, other configs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2u_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_gpu_tests.log
python bench.py > gpurun_out/r2u_bench_lurkfib20.json 2> gpurun_out/r2u_bench_lurkfib20.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2u_bench_ref.json 2> gpurun_out/r2u_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2u_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:merkle_ --csv --log-file gpurun_out/r2u_merkle_dram.csv \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2u_ncu_merkle_dram.log 2>&1
ncu --set full --clock-control none -k regex:lde_k --launch-skip 9 -c 6 -o gpurun_out/r2u_lde \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2u_ncu_lde.log 2>&1
LSK_PHASES=1 python tools/prove_once.py lurk-fib 20 2 > gpurun_out/r2u_phases_lurkfib20.log 2>&1
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2u_bench_lurklcs22.json 2> gpurun_out/r2u_bench_lurklcs22.err
python bench.py --workload lurk-fib --log-rows 22 --shards-per-gpu 4 --steps 2 --no-cpu-baseline > gpurun_out/r2u_bench_lurkfib24.json 2> gpurun_out/r2u_bench_lurkfib24.err
python bench.py --workload lurk-sum --steps 3 --no-cpu-baseline > gpurun_out/r2u_bench_lurksum20.json 2> gpurun_out/r2u_bench_lurksum20.err
python bench.py --workload fib --steps 10 --no-cpu-baseline > gpurun_out/r2u_bench_fib20.json 2> gpurun_out/r2u_bench_fib20.err
python bench.py --workload fib --log-rows 22 --shards-per-gpu 4 --steps 3 --no-cpu-baseline > gpurun_out/r2u_bench_fib24.json 2> gpurun_out/r2u_bench_fib24.err
tail -3 gpurun_out/r2u_gpu_tests.log; for f in lurkfib20 lurklcs22 lurkfib24 lurksum20 fib20 fib24; do head -c 260 gpurun_out/r2u_bench_$f.json | tail -c 150; echo; done
