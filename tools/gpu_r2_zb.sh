#!/bin/bash
# round 2, GPU call ZB: threads per CTA of the compiled permutation-row kernel (LSK_JIT_PTHREADS 256 / 512 / 768); quotient
# at its per-chip default (768 / 256 for the hasher chips)
mkdir -p gpurun_out
for t in 256 512 768; do
  LSK_JIT_PTHREADS=$t python bench.py --no-cpu-baseline > gpurun_out/r2zb_bench_lurkfib20_p$t.json 2> gpurun_out/r2zb_bench_lurkfib20_p$t.err
  LSK_JIT_PTHREADS=$t python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2zb_bench_lurklcs22_p$t.json 2> gpurun_out/r2zb_bench_lurklcs22_p$t.err
done
python bench.py --workload cons --log-rows 20 --steps 3 --no-cpu-baseline > gpurun_out/r2zb_bench_cons20.json 2> gpurun_out/r2zb_bench_cons20.err
for t in 256 512 768; do for f in lurkfib20 lurklcs22; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2zb_bench_${f}_p$t.json') if l.startswith('{')][-1])
print('$f p$t', round(d['ms_per_step'],2), 'perm_rows', round(d['kernels']['perm_rows']['ms_per_step'],2), 'quotient', round(d['kernels']['quotient']['ms_per_step'],2), 'jit', d['config']['compiled_constraint_programs'])
PY
done; done
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2zb_bench_cons20.json') if l.startswith('{')][-1])
print('cons20', round(d['ms_per_step'],2), 'quotient', round(d['kernels']['quotient']['ms_per_step'],2), 'jit', d['config']['compiled_constraint_programs'])
PY
