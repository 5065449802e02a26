#!/bin/bash
# round 2, GPU call P: call O's benches again with the cubin cache rebuilt for the edited field.cuh (O ran interpreted)
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/r2p_bench_lurkfib20.json 2> gpurun_out/r2p_bench_lurkfib20.err
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2p_bench_lurklcs22.json 2> gpurun_out/r2p_bench_lurklcs22.err
python -m pytest tests -m gpu -x -q -k "compiled or side_stream or lurk_bench" > gpurun_out/r2p_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_gpu_tests.log
tail -2 gpurun_out/r2p_gpu_tests.log; for f in lurkfib20 lurklcs22; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2p_bench_$f.json') if l.startswith('{')][-1])
k=d['kernels']
print('$f', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'jit', d['config']['compiled_constraint_programs'], 'leaf', round(k['merkle_leaf']['ms_per_step'],2), 'layer', round(k['merkle_layer']['ms_per_step'],2), 'quot', round(k['quotient']['ms_per_step'],2), d['roofline']['compute']['frac'])
PY
done
