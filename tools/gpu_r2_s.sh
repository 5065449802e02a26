#!/bin/bash
# round 2, GPU call S: opened values leave the device per (round, height) group, tallest first (finer overlap of the host
# transcript with the opening kernels); reduced openings at 4 columns in flight
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "bit_exact or stage or lurk or side_stream or sharded or hasher or u64 or fib" > gpurun_out/r2s_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2s_bench_lurkfib20.json 2> gpurun_out/r2s_bench_lurkfib20.err
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2s_bench_lurklcs22.json 2> gpurun_out/r2s_bench_lurklcs22.err
python bench.py --workload fib --steps 10 --no-cpu-baseline > gpurun_out/r2s_bench_fib20.json 2> gpurun_out/r2s_bench_fib20.err
tail -2 gpurun_out/r2s_gpu_tests.log; for f in lurkfib20 lurklcs22 fib20; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2s_bench_$f.json') if l.startswith('{')][-1])
k=d['kernels']; s=sum(v['ms_per_step'] for v in k.values())
print('$f', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'kernel sum', round(s,2))
PY
done
