#!/bin/bash
# round 2, GPU call AD: last check of the artefacts as they stand (library and cubin cache rebuilt from the committed sources)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ad_smoke.log 2>&1; tail -1 gpurun_out/r2ad_smoke.log
python -m pytest tests -m gpu -x -q -k "compiled or lurk_bench or side_stream or fib_proof" > gpurun_out/r2ad_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ad_gpu_tests.log
python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2ad_bench_lurkfib20.json 2> gpurun_out/r2ad_bench_lurkfib20.err
tail -2 gpurun_out/r2ad_gpu_tests.log; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2ad_bench_lurkfib20.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'jit', d['config']['compiled_constraint_programs'])
PY
