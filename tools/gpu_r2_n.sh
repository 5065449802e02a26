#!/bin/bash
# round 2, GPU calls N / N2: pipe placement of the plain additions of the width-16 permutation (tools/perm_variants.cu; field.cuh
# LSK_ALU_ADD: bit 0 every fadd, bit 2 mont_reduce's subtraction, bit 3 round-constant additions, bit 4 the column sums of
# the external layer, bit 5 the M4 additions, bit 6 the internal layer's "+ sum")
mkdir -p gpurun_out
for f in tools/_pv/perm_variant_*; do $f; done | tee gpurun_out/r2n_perm_variants2.jsonl
