#!/bin/bash
# round 2, GPU call ZA: compiled quotient at 768 threads per CTA as the default: parity, the hasher-chip workload (cons),
# the Lurk benches
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "compiled or hasher or cons or lurk or bit_exact or stage or u64 or side_stream" > gpurun_out/r2za_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2za_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2za_bench_lurkfib20.json 2> gpurun_out/r2za_bench_lurkfib20.err
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2za_bench_lurklcs22.json 2> gpurun_out/r2za_bench_lurklcs22.err
python bench.py --workload cons --log-rows 20 --steps 3 --no-cpu-baseline > gpurun_out/r2za_bench_cons20.json 2> gpurun_out/r2za_bench_cons20.err
LSK_JIT_QTHREADS=256 LSK_JIT=1 python bench.py --workload cons --log-rows 20 --steps 3 --no-cpu-baseline > gpurun_out/r2za_bench_cons20_q256.json 2> gpurun_out/r2za_bench_cons20_q256.err
tail -2 gpurun_out/r2za_gpu_tests.log; for f in lurkfib20 lurklcs22 cons20 cons20_q256; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2za_bench_$f.json') if l.startswith('{')][-1])
print('$f', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'quotient', round(d['kernels']['quotient']['ms_per_step'],2), 'jit', d['config']['compiled_constraint_programs'])
PY
done
