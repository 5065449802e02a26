#!/bin/bash
# round 2, GPU call F: lazy 64-bit accumulation in the LogUp sums / extension products, grouped inversions in the compiled perm rows
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "compiled or lurk or stage or imported or bit_exact or hasher or u64" > gpurun_out/r2f_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2f_bench_lurkfib20.json 2> gpurun_out/r2f_bench_lurkfib20.err
ncu --set full --clock-control none --import-source on -k regex:lsk_quotient_jit --launch-skip 1 -c 1 -o gpurun_out/r2f_quot_builtin \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2f_ncu_quot.log 2>&1
tail -3 gpurun_out/r2f_gpu_tests.log; head -c 300 gpurun_out/r2f_bench_lurkfib20.json
