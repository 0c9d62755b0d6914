#!/bin/bash
mkdir -p gpurun_out
for pr in tf32 bf16; do for ki in 1 2 4; do for c in 64 200; do echo -n "C=$c $pr KI=$ki: "; ( LIFT_C=$c OCCD_LIFT_KI=$ki OCCDEPTH_PRECISION=$pr timeout 100 python tools/lift_bench.py ) 2>&1 | tail -1; done; done; done
( timeout 600 python -m pytest tests/test_gpu_sfa.py tests/test_gpu_golden.py -q -m gpu -x ) 2>&1 | tail -3
