#!/bin/bash
# round 2, GPU call AB (2 GPUs): the final code in a device group, once more after the last kernel changes -- 2-GPU parity test, the driver's N=2 launch with its
# strong sub-record, memcheck of a small Lurk prove on one GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "two_gpus or device_group or side_stream" > gpurun_out/r2ab_gpu_tests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ab_gpu_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r2ab_bench_lurkfib20_2gpu.json 2> gpurun_out/r2ab_bench_lurkfib20_2gpu.err
tail -3 gpurun_out/r2ab_gpu_tests_2gpu.log; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2ab_bench_lurkfib20_2gpu.json') if l.startswith('{')][-1])
print(d['value']/1e6, d['ms_per_step'], d['e2e']); print(d['strong'])
PY
