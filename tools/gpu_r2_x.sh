#!/bin/bash
# round 2, GPU call X: reduced-opening targets y computed on the device (reduce_y_kernel)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "merkle or bit_exact or lurk or stage or sharded or hasher or u64 or side_stream or fib or cons or zstore" > gpurun_out/r2x_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2x_bench_lurkfib20.json 2> gpurun_out/r2x_bench_lurkfib20.err
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2x_bench_lurklcs22.json 2> gpurun_out/r2x_bench_lurklcs22.err
python bench.py --workload fib --steps 10 --no-cpu-baseline > gpurun_out/r2x_bench_fib20.json 2> gpurun_out/r2x_bench_fib20.err
tail -3 gpurun_out/r2x_gpu_tests.log; for f in lurkfib20 lurklcs22 fib20; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2x_bench_$f.json') if l.startswith('{')][-1])
k=d['kernels']
print('$f', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'top', round(k['merkle_top']['ms_per_step'],3), k['merkle_top']['launches_per_step'])
PY
done
