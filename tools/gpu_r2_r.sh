#!/bin/bash
# round 2, GPU call R: reduced-openings kernel with 2 / 4 / 8 columns in flight (LSK_RO_COLS), parity of the changed path
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "bit_exact or stage or lurk_bench or side_stream" > gpurun_out/r2r_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_gpu_tests.log
for c in 8 4 2; do LSK_RO_COLS=$c python bench.py --no-cpu-baseline > gpurun_out/r2r_bench_lurkfib20_ro$c.json 2> gpurun_out/r2r_bench_lurkfib20_ro$c.err; done
LSK_RO_COLS=8 python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2r_bench_lurklcs22_ro8.json 2>/dev/null
LSK_RO_COLS=2 python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2r_bench_lurklcs22_ro2.json 2>/dev/null
tail -2 gpurun_out/r2r_gpu_tests.log; for f in lurkfib20_ro8 lurkfib20_ro4 lurkfib20_ro2 lurklcs22_ro8 lurklcs22_ro2; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2r_bench_$f.json') if l.startswith('{')][-1])
k=d['kernels']
print('$f', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'reduce', round(k['reduce_openings']['ms_per_step'],3), 'open', round(k['open']['ms_per_step'],3))
PY
done
