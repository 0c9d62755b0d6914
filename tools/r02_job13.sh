#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -x ) 2>&1 | tail -3
for pr in tf32 bf16; do for s in head_c32_d1 head_c32_d3 up1_conv2 up2_conv2 up4_conv2 up16_conv1 x_c64_n64 l1_k113_16 b1_expand proj_94; do
  echo -n "$pr "; ( OCCDEPTH_PRECISION=$pr timeout 100 python tools/conv_bench.py $s ) 2>&1 | tail -1 | cut -c1-110
done; done
for pr in tf32 bf16; do
  echo "$pr"; ( timeout 300 python bench.py --precision $pr --steps 10 --warmup 3 --no-cpu --no-modes ) 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"
done
