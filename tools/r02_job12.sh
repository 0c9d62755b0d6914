#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -x ) 2>&1 | tail -3
for dbg in 0 3; do for pr in tf32 bf16; do for s in b1_expand b1_expand_relu proj_94 bneck5_16_64 up2_conv2; do
  echo -n "dbg=$dbg $pr "; ( OCCD_DEBUG_EPI=$dbg OCCDEPTH_PRECISION=$pr timeout 100 python tools/conv_bench.py $s ) 2>&1 | tail -1 | cut -c1-110
done; done; done
for dbg in 0 3; do for pr in tf32 bf16; do
  echo "dbg=$dbg $pr"; ( OCCD_DEBUG_EPI=$dbg timeout 300 python bench.py --precision $pr --steps 10 --warmup 3 --no-cpu --no-modes ) 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"
done; done
