#!/bin/bash
# round 2, GPU call O: Merkle / Poseidon2 kernels with mont_reduce's subtraction on the ALU pipe (LSK_ALU_ADD=4), the
# side-stream parity test
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "merkle or poseidon or lurk or bit_exact or side_stream or fri or stage or sharded or verif" > gpurun_out/r2o_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2o_bench_lurkfib20.json 2> gpurun_out/r2o_bench_lurkfib20.err
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2o_bench_lurklcs22.json 2> gpurun_out/r2o_bench_lurklcs22.err
python tools/poseidon2_sweep.py --min-log 22 --max-log 24 > gpurun_out/r2o_sweep.log 2>&1
tail -3 gpurun_out/r2o_gpu_tests.log; for f in lurkfib20 lurklcs22; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2o_bench_$f.json') if l.startswith('{')][-1])
k=d['kernels']
print('$f', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'leaf', round(k['merkle_leaf']['ms_per_step'],2), 'layer', round(k['merkle_layer']['ms_per_step'],2), 'top', round(k['merkle_top']['ms_per_step'],2), d['roofline']['compute']['frac'])
PY
done; grep '"log_n": 24' gpurun_out/r2o_sweep.log | cut -c1-160
