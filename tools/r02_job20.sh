#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dropin.py -q -m gpu -x ) 2>&1 | tail -3
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-modes --dump-profile ) 2> gpurun_out/r02d_bench.err | tail -1 > gpurun_out/r02d_bench_tf32.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d_bench_tf32.json'))
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'seq', d['e2e']['sequential_ms_per_step'], 'pipe', d['e2e']['pipelined_ms_per_step'], d['e2e']['pipelined_error'], 'frac', d['roofline']['frac'], 'lift', d['roofline_lift']['frac'])
PY
ls gpurun_out | grep -i profile
