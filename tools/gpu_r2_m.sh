#!/bin/bash
# round 2, GPU call M: record upload in waves on the side stream (trace generation starts while later tables are still
# crossing PCIe): full GPU suite, default bench (e2e is what should move), A/B with LSK_NO_SIDE_STREAM=1, other workloads
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2m_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2m_bench_lurkfib20.json 2> gpurun_out/r2m_bench_lurkfib20.err
LSK_NO_SIDE_STREAM=1 python bench.py --no-cpu-baseline > gpurun_out/r2m_bench_lurkfib20_noside.json 2> gpurun_out/r2m_bench_lurkfib20_noside.err
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2m_bench_lurklcs22.json 2> gpurun_out/r2m_bench_lurklcs22.err
python bench.py --workload fib --steps 10 --no-cpu-baseline > gpurun_out/r2m_bench_fib20.json 2> gpurun_out/r2m_bench_fib20.err
python bench.py --workload lurk-fib --log-rows 22 --shards-per-gpu 4 --steps 2 --no-cpu-baseline > gpurun_out/r2m_bench_lurkfib24.json 2> gpurun_out/r2m_bench_lurkfib24.err
compute-sanitizer --tool racecheck --print-limit 5 python tools/prove_once.py fib 12 1 > gpurun_out/r2m_racecheck.log 2>&1; tail -3 gpurun_out/r2m_racecheck.log
tail -3 gpurun_out/r2m_gpu_tests.log; for f in lurkfib20 lurkfib20_noside lurklcs22 fib20 lurkfib24; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2m_bench_$f.json') if l.startswith('{')][-1])
print('$f', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2))
PY
done
