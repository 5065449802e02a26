#!/bin/bash
# round 2, GPU call AF2: ncu captures of the final compiled quotient kernels (768 threads per CTA) of the tall chips of a Lurk
# fib prove (the tiny chips launch first, on the side stream: skipped), for the stall breakdown quoted in DESIGN.md
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:lsk_quotient_jit --launch-skip 13 -c 5 -o gpurun_out/r2af_quot_tall \
    python tools/prove_once.py lurk-fib 20 1 > gpurun_out/r2af_ncu_quot.log 2>&1
tail -2 gpurun_out/r2af_ncu_quot.log
