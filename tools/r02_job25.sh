#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_net2d.py -q -m gpu -x ) 2>&1 | tail -3
( OCCDEPTH_PRECISION=tf32 timeout 300 python tools/dw_bench.py ) 2>&1 | grep "pool=1"
