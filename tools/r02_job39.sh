#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -x ) 2>&1 | tail -4
for ring in 1 0; do for s in head_c32_d1 head_c32_d2 head_c32_d3 head_c32_n2; do
  echo -n "ring=$ring tf32 "; ( OCCD_HALOX_RING=$ring OCCDEPTH_PRECISION=tf32 timeout 100 python tools/conv_bench.py $s ) 2>&1 | tail -1 | cut -c1-120
done; done
