#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet3d.py tests/test_gpu_golden.py tests/test_gpu_slab.py tests/test_gpu_config2.py -q -m gpu -x ) 2>&1 | tail -3
for ring in 1 0; do echo "ring=$ring"; ( OCCD_HALOX_RING=$ring timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-modes ) 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['step_ms']['timed_region_repetitions_ms_per_step'], d['roofline']['frac'])"; done
