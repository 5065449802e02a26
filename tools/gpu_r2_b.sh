#!/bin/bash
# round 2, GPU call B: compiled interactions (perm rows + quotient LogUp part), L2-resident LDE column chunks, JIT cache fix
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_gpu_tests.log
LSK_JIT_DEBUG=1 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench_lurkfib20.json 2> gpurun_out/r2b_bench_lurkfib20.err
LSK_LDE_L2_MB=0 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench_lurkfib20_nochunk.json 2> /dev/null
LSK_PHASES=1 python tools/prove_once.py lurk-fib 20 2 > gpurun_out/r2b_phases_lurkfib20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsk_quotient_jit --launch-skip 8 -c 4 -o gpurun_out/r2b_quot_jit \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2b_ncu_quot.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsk_perm_rows_jit --launch-skip 8 -c 4 -o gpurun_out/r2b_perm_jit \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2b_ncu_perm.log 2>&1
ncu --set full --clock-control none -k regex:lde_k --launch-skip 9 -c 9 -o gpurun_out/r2b_lde \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2b_ncu_lde.log 2>&1
tail -3 gpurun_out/r2b_gpu_tests.log; head -c 600 gpurun_out/r2b_bench_lurkfib20.json; grep -c miss gpurun_out/r2b_bench_lurkfib20.err
