#!/bin/bash
# round 2, GPU call J: K2 instantiated per size (register budget A/B), two-rows-per-thread compiled quotient (A/B),
# tiny chips on a side stream (A/B),
# DRAM bytes of every Merkle launch of one prove (the `traffic` figure of the dominant kernel family)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "lde or coset or lurk or compiled or bit_exact or stage or sharded or hasher or u64" > gpurun_out/r2j_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2j_bench_lurkfib20.json 2> gpurun_out/r2j_bench_lurkfib20.err
LSK_K2_MINB=2 python bench.py --no-cpu-baseline > gpurun_out/r2j_bench_lurkfib20_k2minb2.json 2> gpurun_out/r2j_bench_lurkfib20_k2minb2.err
LSK_JIT_ROWS=1 python bench.py --no-cpu-baseline > gpurun_out/r2j_bench_lurkfib20_rows1.json 2> gpurun_out/r2j_bench_lurkfib20_rows1.err
LSK_NO_SIDE_STREAM=1 python bench.py --no-cpu-baseline > gpurun_out/r2j_bench_lurkfib20_noside.json 2> gpurun_out/r2j_bench_lurkfib20_noside.err
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2j_bench_lurklcs22.json 2> gpurun_out/r2j_bench_lurklcs22.err
LSK_K2_MINB=2 python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2j_bench_lurklcs22_k2minb2.json 2> /dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:merkle_ --csv --log-file gpurun_out/r2j_merkle_dram.csv \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2j_ncu_merkle_dram.log 2>&1
tail -3 gpurun_out/r2j_gpu_tests.log; for f in lurkfib20 lurkfib20_k2minb2 lurkfib20_rows1 lurkfib20_noside lurklcs22 lurklcs22_k2minb2; do head -c 230 gpurun_out/r2j_bench_$f.json | tail -c 120; echo; done
