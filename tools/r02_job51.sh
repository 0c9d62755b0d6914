#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( timeout 2400 python -m pytest tests -q -m gpu -x ) > gpurun_out/r02_final_gpu_tests.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_final_gpu_tests.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -1
( timeout 900 python bench.py --dump-profile ) > gpurun_out/r02l_bench_tf32.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/r02l_bench_tf32.json)"
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_tf32.json'))
print('ms', d['ms_per_step'], d['step_ms']['timed_region_repetitions_ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['pipelined_error'], 'frac', d['roofline']['frac'], 'lift', d['roofline_lift']['frac'])
print('tm', d['throughput_mode']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d['cuda_reference']['tf32_allowed']['ms_per_step'], d['cuda_reference']['fp32_only']['ms_per_step'])
PY
