#!/bin/bash
# round 2, GPU call AC: read-only (LDG.CONSTANT) loads of the cells in the openings kernel and of the permutation columns in
# the compiled quotient
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "compiled or lurk_bench or bit_exact or stage or side_stream" > gpurun_out/r2ac_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ac_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2ac_bench_lurkfib20.json 2> gpurun_out/r2ac_bench_lurkfib20.err
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2ac_bench_lurklcs22.json 2> gpurun_out/r2ac_bench_lurklcs22.err
tail -2 gpurun_out/r2ac_gpu_tests.log; for f in lurkfib20 lurklcs22; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2ac_bench_$f.json') if l.startswith('{')][-1])
k=d['kernels']
print('$f', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'open', round(k['open']['ms_per_step'],2), 'quotient', round(k['quotient']['ms_per_step'],2), 'jit', d['config']['compiled_constraint_programs'])
PY
done
