#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python bench.py ) 2> gpurun_out/r02f_bench.err | tail -1 > gpurun_out/r02f_bench_tf32.json; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench_tf32.json'))
print('steps', d['steps'], d['warmup'], 'ms', d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'lift', d['roofline_lift']['frac'])
PY
( timeout 900 python bench.py --steps 20 --warmup 5 ) 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('20/5 full', d['ms_per_step'], d['step_ms'], d['e2e']['ms_per_step'])"
