#!/bin/bash
# round 2, GPU call Z: threads per CTA of the compiled quotient kernel (LSK_JIT_QTHREADS 256 / 512 / 768: instruction
# streams per SM 3 / 1-2 / 1)
mkdir -p gpurun_out
for t in 256 512 768; do
  LSK_JIT_QTHREADS=$t python bench.py --no-cpu-baseline > gpurun_out/r2z_bench_lurkfib20_q$t.json 2> gpurun_out/r2z_bench_lurkfib20_q$t.err
  LSK_JIT_QTHREADS=$t python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2z_bench_lurklcs22_q$t.json 2> gpurun_out/r2z_bench_lurklcs22_q$t.err
done
LSK_JIT_QTHREADS=512 python -m pytest tests -m gpu -x -q -k "lurk_bench or compiled" > gpurun_out/r2z_gpu_tests_q512.log 2>&1; tail -1 gpurun_out/r2z_gpu_tests_q512.log
for t in 256 512 768; do for f in lurkfib20 lurklcs22; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2z_bench_${f}_q$t.json') if l.startswith('{')][-1])
print('$f q$t', round(d['ms_per_step'],2), 'quotient', round(d['kernels']['quotient']['ms_per_step'],2), 'jit', d['config']['compiled_constraint_programs'])
PY
done; done
