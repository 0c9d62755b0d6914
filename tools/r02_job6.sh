#!/bin/bash
mkdir -p gpurun_out
for pr in tf32 bf16; do for s in b1_expand bneck5_16_64 up2_conv2 proj_94; do ( OCCDEPTH_PRECISION=$pr timeout 120 python tools/conv_trace.py $s ) > gpurun_out/r02_trace2_${pr}_$s.txt 2>&1; echo "== $pr $s"; head -9 gpurun_out/r02_trace2_${pr}_$s.txt; done; done
