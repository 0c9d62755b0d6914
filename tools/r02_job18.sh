#!/bin/bash
mkdir -p gpurun_out
for pr in tf32 bf16; do for nv in 4 2; do for spp in 1 2 4; do echo -n "C=64 $pr NV=$nv SPP=$spp: "; ( OCCD_LIFT_NV=$nv OCCD_LIFT_SPP=$spp OCCDEPTH_PRECISION=$pr timeout 100 python tools/lift_bench.py ) 2>&1 | tail -1; done; done; done
for spp in 1 2 4; do ( OCCD_LIFT_SPP=$spp timeout 600 python -m pytest tests/test_gpu_sfa.py -q -m gpu -x ) 2>&1 | tail -1; done
