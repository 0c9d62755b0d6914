#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_ops.py tests/test_gpu_net2d.py tests/test_gpu_golden.py -q -m gpu -x -k "bf16" ) 2>&1 | tail -2
( timeout 600 python -m pytest tests/test_gpu_config2.py -q -m gpu -x ) 2>&1 | tail -2
( timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu --no-modes ) 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bf16', d['ms_per_step'], d['step_ms']['timed_region_repetitions_ms_per_step'])"
python -c "
import json
for f in ('config2_parity_bf16','config2_flospdepth_parity_bf16'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ssc_logit']['rel'], d['occ_logit']['rel'], d['argmax_agreement'])"
