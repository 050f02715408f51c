#!/bin/bash
# one gpurun call: cigar_scan variants (chunks per warp per run) on the C3-shaped haploid batch
for g in 4 8 16; do
  lib=svim_asm_b200/libtune_g$g.so
  [ $g = 4 ] && lib=svim_asm_b200/libsvimasm_b200.so
  echo "== G=$g"
  SVIM_ASM_B200_LIB=$PWD/$lib timeout 200 python tools/perf_scan.py --scale ${1:-1.0} --iters 10 2>&1 | tail -2
done
