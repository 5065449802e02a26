#!/bin/bash
# round 2, GPU call H: the evidence set -- full GPU suite, the default bench with its CPU leg, the reference arm, the launch
# list of the bench command, one full ncu capture of the top kernel (mixed-height Merkle layers), the other BASELINE
# configs (lcs 2^22, fib 2^24 on one GPU), ZStore batch hashing crossover, Poseidon2 sweep to 2^28 on one GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_gpu_tests.log
python bench.py > gpurun_out/r2h_bench_lurkfib20.json 2> gpurun_out/r2h_bench_lurkfib20.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2h_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:merkle_flat_kernel --launch-skip 5 -c 4 -o gpurun_out/r2h_merkle_flat \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2h_ncu_merkle.log 2>&1
LSK_PHASES=1 python tools/prove_once.py lurk-fib 20 2 > gpurun_out/r2h_phases_lurkfib20.log 2>&1
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2h_bench_lurklcs22.json 2> gpurun_out/r2h_bench_lurklcs22.err
python bench.py --workload lurk-fib --log-rows 22 --shards-per-gpu 4 --steps 2 --no-cpu-baseline > gpurun_out/r2h_bench_lurkfib24.json 2> gpurun_out/r2h_bench_lurkfib24.err
python bench.py --workload lurk-sum --steps 3 --no-cpu-baseline > gpurun_out/r2h_bench_lurksum20.json 2> gpurun_out/r2h_bench_lurksum20.err
python bench.py --workload fib --steps 10 --no-cpu-baseline > gpurun_out/r2h_bench_fib20.json 2> gpurun_out/r2h_bench_fib20.err
python tools/zstore_batch_bench.py --out gpurun_out/r2h_zstore_batch.jsonl > gpurun_out/r2h_zstore_batch.log 2>&1
python tools/poseidon2_sweep.py --min-log 24 --max-log 28 --out gpurun_out/r2h_poseidon2_sweep_1gpu.jsonl > gpurun_out/r2h_sweep1.log 2>&1
tail -3 gpurun_out/r2h_gpu_tests.log; for f in lurkfib20 lurklcs22 lurkfib24 lurksum20 fib20; do head -c 260 gpurun_out/r2h_bench_$f.json; echo; done; tail -3 gpurun_out/r2h_zstore_batch.log | cut -c1-300; tail -2 gpurun_out/r2h_sweep1.log | cut -c1-200
