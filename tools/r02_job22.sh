#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet3d.py tests/test_gpu_slab.py tests/test_gpu_golden.py -q -m gpu -x ) 2>&1 | tail -4
for g in 1 0; do echo "grouped=$g"; ( OCCDEPTH_CONVT_GROUPED=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-modes --dump-profile ) 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"; cp gpurun_out/plan_profile_tf32.json gpurun_out/plan_profile_tf32_g$g.json; done
