#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_net2d.py -q -m gpu -x ) 2>&1 | tail -3
for px in 8 4; do echo "px=$px"; ( OCCD_DW_PX=$px OCCDEPTH_PRECISION=tf32 timeout 300 python tools/dw_bench.py ) 2>&1 | grep "pool=1"; done
for px in 8 4; do echo "px=$px"; ( OCCD_DW_PX=$px timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-modes ) 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"; done
