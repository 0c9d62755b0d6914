#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-modes ) 2> gpurun_out/r02c_bench.err | tail -1 > gpurun_out/r02c_bench_tf32.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench_tf32.json'))
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'seq', d['e2e']['sequential_ms_per_step'], 'pipe', d['e2e']['pipelined_ms_per_step'], d['e2e']['pipelined_error'], 'frac', d['roofline']['frac'], 'lift', d['roofline_lift']['frac'])
PY
for r in 8 12; do echo "epi_rate=$r"; ( OCCD_EPI_RATE=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-modes ) 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"; done
