#!/bin/bash
# round 2, GPU call N: pipe placement of the plain additions inside the width-16 permutation (tools/perm_variants.cu,
# field.cuh LSK_ALU_ADD bit 0 fadd, bit 1 fsub, bit 2 mont_reduce's high-word subtraction)
mkdir -p gpurun_out
for k in 0 1 2 3 4 5 6 7 0; do tools/_pv/perm_variant_$k; done | tee gpurun_out/r2n_perm_variants.jsonl
