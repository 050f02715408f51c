#!/bin/bash
# one gpurun call: cigar_scan tuning builds on the C3-shaped haploid batch
for tag in ${TAGS:-default}; do
  lib=svim_asm_b200/libtune_$tag.so
  [ $tag = default ] && lib=svim_asm_b200/libsvimasm_b200.so
  echo "== $tag"
  SVIM_ASM_B200_LIB=$PWD/$lib timeout 200 python tools/perf_scan.py --scale ${1:-1.0} --iters 10 2>&1 | tail -2 | cut -c1-140
done
