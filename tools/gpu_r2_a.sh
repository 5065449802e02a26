#!/bin/bash
# round 2, GPU call A: the GPU test suite, the new default bench, phase times and launch list of the Lurk fib workload,
# ncu captures of the kernels VERDICT r1 names (mixed-height Merkle layers, compiled quotient, LDE K2)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_gpu_tests.log
python bench.py > gpurun_out/r2a_bench_lurkfib20.json 2> gpurun_out/r2a_bench_lurkfib20.err
python bench.py --impl reference > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
LSK_PHASES=1 python tools/prove_once.py lurk-fib 20 2 > gpurun_out/r2a_phases_lurkfib20.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_lurkfib20.csv \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2a_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:merkle_flat_kernel -c 8 -o gpurun_out/r2a_merkle_flat \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2a_ncu_merkle.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsk_quotient_jit --launch-skip 30 -c 4 -o gpurun_out/r2a_quot_jit \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2a_ncu_quot.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lde_k -c 9 -o gpurun_out/r2a_lde \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2a_ncu_lde.log 2>&1
tail -3 gpurun_out/r2a_gpu_tests.log; head -c 1500 gpurun_out/r2a_bench_lurkfib20.json
