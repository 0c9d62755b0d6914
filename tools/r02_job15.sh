#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_sfa.py tests/test_gpu_golden.py -q -m gpu -x ) 2>&1 | tail -3
for c in 64 200; do for pr in tf32 bf16; do echo -n "C=$c $pr: "; ( LIFT_C=$c OCCDEPTH_PRECISION=$pr timeout 100 python tools/lift_bench.py ) 2>&1 | tail -1; done; done
