#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -x ) 2>&1 | tail -2
for ring in 0 1; do for dbg in 0 4; do for s in head_c32_d1 head_c32_d3; do
  echo -n "ring=$ring dbg=$dbg tf32 "; ( OCCD_HALOX_RING=$ring OCCD_DEBUG_EPI=$dbg OCCDEPTH_PRECISION=tf32 timeout 100 python tools/conv_bench.py $s ) 2>&1 | tail -1 | cut -c1-100
done; done; done
for dbg in 0 4; do echo -n "bf16 dbg=$dbg "; ( OCCD_DEBUG_EPI=$dbg OCCDEPTH_PRECISION=bf16 timeout 100 python tools/conv_bench.py head_c32_d1 ) 2>&1 | tail -1 | cut -c1-100; done
