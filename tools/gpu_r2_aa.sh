#!/bin/bash
# round 2, GPU call AA: full GPU suite and the bench configs on the final code (compiled quotient and permutation rows at 768 threads
# per CTA)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_gpu_tests.log
python bench.py > gpurun_out/r2aa_bench_lurkfib20.json 2> gpurun_out/r2aa_bench_lurkfib20.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2aa_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2aa_ncu_launches.log 2>&1
python bench.py --workload lurk-lcs --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/r2aa_bench_lurklcs22.json 2> gpurun_out/r2aa_bench_lurklcs22.err
python bench.py --workload lurk-fib --log-rows 22 --shards-per-gpu 4 --steps 2 --no-cpu-baseline > gpurun_out/r2aa_bench_lurkfib24.json 2> gpurun_out/r2aa_bench_lurkfib24.err
python bench.py --workload lurk-sum --steps 3 --no-cpu-baseline > gpurun_out/r2aa_bench_lurksum20.json 2> gpurun_out/r2aa_bench_lurksum20.err
python bench.py --workload fib --steps 10 --no-cpu-baseline > gpurun_out/r2aa_bench_fib20.json 2> gpurun_out/r2aa_bench_fib20.err
python bench.py --workload fib --log-rows 22 --shards-per-gpu 4 --steps 3 --no-cpu-baseline > gpurun_out/r2aa_bench_fib24.json 2> gpurun_out/r2aa_bench_fib24.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2aa_smoke.log 2>&1; tail -1 gpurun_out/r2aa_smoke.log
tail -3 gpurun_out/r2aa_gpu_tests.log; for f in lurkfib20 lurklcs22 lurkfib24 lurksum20 fib20 fib24; do python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2aa_bench_$f.json') if l.startswith('{')][-1])
print('$f', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['config'].get('parity_vs_oracle'))
PY
done
