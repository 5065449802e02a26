#!/bin/bash
# round 2, GPU call ZC: compiled permutation-row kernel at 768 threads per CTA as the default (parity), 1024 as a variant
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "compiled or hasher or cons or lurk or bit_exact or stage or u64 or side_stream or two_gpus" > gpurun_out/r2zc_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2zc_gpu_tests.log
for t in 768 1024; do
  LSK_JIT_PTHREADS=$t python bench.py --no-cpu-baseline > gpurun_out/r2zc_bench_lurkfib20_p$t.json 2> gpurun_out/r2zc_bench_lurkfib20_p$t.err
  LSK_JIT_PTHREADS=$t python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2zc_bench_lurklcs22_p$t.json 2> gpurun_out/r2zc_bench_lurklcs22_p$t.err
done
tail -2 gpurun_out/r2zc_gpu_tests.log
for t in 768 1024; do for f in lurkfib20 lurklcs22; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2zc_bench_${f}_p$t.json') if l.startswith('{')][-1])
print('$f p$t', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'perm_rows', round(d['kernels']['perm_rows']['ms_per_step'],2), 'jit', d['config']['compiled_constraint_programs'])
PY
done; done
