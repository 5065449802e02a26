#!/bin/bash
# round 2, GPU call V (2 GPUs): openings of a device group through the same round pipeline as one GPU (all-reduce per round,
# transcript overlapped); single-GPU parity of the unified path
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "two_gpus or device_group or bit_exact or lurk_bench or stage" > gpurun_out/r2v_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_gpu_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r2v_bench_lurkfib20_2gpu.json 2> gpurun_out/r2v_bench_lurkfib20_2gpu.err
python bench.py --no-cpu-baseline > gpurun_out/r2v_bench_lurkfib20.json 2> gpurun_out/r2v_bench_lurkfib20.err
tail -3 gpurun_out/r2v_gpu_tests.log; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2v_bench_lurkfib20_2gpu.json') if l.startswith('{')][-1])
print(d['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step']); print(d['strong'])
d=json.loads([l for l in open('gpurun_out/r2v_bench_lurkfib20.json') if l.startswith('{')][-1])
print('1gpu', d['ms_per_step'], d['e2e']['ms_per_step'])
PY
