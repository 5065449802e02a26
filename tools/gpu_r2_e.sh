#!/bin/bash
# round 2, GPU call E: read-only (LDG.CONSTANT) trace loads in the compiled programs, bulk-async K3 variant
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "compiled or lurk or lde or stage or imported" > gpurun_out/r2e_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2e_bench_lurkfib20.json 2> gpurun_out/r2e_bench_lurkfib20.err
LSK_LDE_BULK=1 python bench.py --no-cpu-baseline > gpurun_out/r2e_bench_lurkfib20_bulk.json 2> gpurun_out/r2e_bench_lurkfib20_bulk.err
LSK_LDE_BULK=1 python -m pytest tests -m gpu -x -q -k "lde or fib_proof_is_bit_exact" > gpurun_out/r2e_gpu_tests_bulk.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_gpu_tests_bulk.log
ncu --set full --clock-control none --import-source on -k regex:lsk_quotient_jit --launch-skip 1 -c 1 -o gpurun_out/r2e_quot_builtin \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2e_ncu_quot.log 2>&1
LSK_LDE_BULK=1 ncu --set full --clock-control none -k regex:lde_k3 --launch-skip 3 -c 2 -o gpurun_out/r2e_k3_bulk \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2e_ncu_k3b.log 2>&1
ncu --set full --clock-control none -k regex:lde_k3 --launch-skip 3 -c 2 -o gpurun_out/r2e_k3_plain \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2e_ncu_k3.log 2>&1
tail -3 gpurun_out/r2e_gpu_tests.log; tail -3 gpurun_out/r2e_gpu_tests_bulk.log; head -c 300 gpurun_out/r2e_bench_lurkfib20.json; echo; head -c 300 gpurun_out/r2e_bench_lurkfib20_bulk.json
