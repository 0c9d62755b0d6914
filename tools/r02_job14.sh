#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 3; do for pr in tf32 bf16; do for s in b1_expand b1_expand_noact expand_188 proj_94 bneck5_16_64 l1_k1_64_16; do
  echo -n "dbg=$dbg $pr "; ( OCCD_DEBUG_EPI=$dbg OCCDEPTH_PRECISION=$pr timeout 100 python tools/conv_bench.py $s ) 2>&1 | tail -1 | cut -c1-110
done; done; done
