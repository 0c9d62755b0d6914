#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2; do for s in b1_expand b1_expand_relu b1_expand_noact proj_94 bneck5_16_64 up2_conv2; do
  echo -n "dbg=$dbg "; ( OCCD_DEBUG_EPI=$dbg OCCDEPTH_PRECISION=tf32 timeout 100 python tools/conv_bench.py $s ) 2>&1 | tail -1 | cut -c1-110
done; done
for c in 64 200; do for pr in tf32 bf16; do for nv in 4 2; do echo -n "C=$c $pr NV=$nv: "; ( LIFT_C=$c OCCD_LIFT_NV=$nv OCCDEPTH_PRECISION=$pr timeout 100 python tools/lift_bench.py ) 2>&1 | tail -1; done; done; done
( timeout 300 python -m pytest tests/test_gpu_sfa.py tests/test_gpu_golden.py -q -m gpu ) 2>&1 | tail -3
