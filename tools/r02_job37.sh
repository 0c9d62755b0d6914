#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( timeout 2400 python -m pytest tests -q -m gpu -x ) > gpurun_out/r02_final_gpu_tests.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_final_gpu_tests.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -1
( timeout 900 python bench.py ) > gpurun_out/r02i_bench_tf32.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/r02i_bench_tf32.json)"
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02i_bench_ref.json 2> gpurun_out/r02i_bench_ref.err; echo "ref rc=$? lines=$(wc -l < gpurun_out/r02i_bench_ref.json)"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i_bench_tf32.json'))
print('ms', d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['pipelined_error'], 'frac', d['roofline']['frac'], d['roofline'].get('traffic'), 'lift', d['roofline_lift']['frac'], 'launches', d['gpu_launches'])
print('cpu', d.get('cpu_baseline')); print('tm', json.dumps(d.get('throughput_mode'))[:160]); print(d['clocks'])
r=json.load(open('gpurun_out/r02i_bench_ref.json')); print('ref', r['value'], r['ms_per_step'])
PY
