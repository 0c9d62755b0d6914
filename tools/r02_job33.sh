#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python bench.py ) 2> gpurun_out/r02g_bench.err | tail -1 > gpurun_out/r02g_bench_tf32.json; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench_tf32.json'))
print('steps', d['steps'], d['warmup'], 'ms', d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'lift', d['roofline_lift']['frac'])
print('cpu', d.get('cpu_baseline')); print('tm', json.dumps(d.get('throughput_mode'))[:200])
PY
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) 2> gpurun_out/r02g_bench_ref.err | tail -1 > gpurun_out/r02g_bench_ref.json; echo "ref rc=$?"; python -c "
import json; r=json.load(open('gpurun_out/r02g_bench_ref.json')); print('ref', r['value'], r['ms_per_step'], r['cpu_baseline'])"
