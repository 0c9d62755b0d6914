#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for pr in tf32 bf16; do
  ( timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02e_ncu_$pr.csv python tools/ncu_forward.py $pr ) > gpurun_out/r02e_ncu_$pr.log 2>&1; echo "ncu $pr rc=$?"; tail -1 gpurun_out/r02e_ncu_$pr.log
done
( timeout 900 python bench.py --steps 20 --warmup 5 --dump-profile ) 2> gpurun_out/r02e_bench.err | tail -1 > gpurun_out/r02e_bench_tf32.json; echo "bench rc=$?"
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) 2> gpurun_out/r02e_bench_ref.err | tail -1 > gpurun_out/r02e_bench_ref.json; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench_tf32.json'))
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'seq', d['e2e']['sequential_ms_per_step'], d['e2e']['pipelined_error'], 'frac', d['roofline']['frac'], 'lift', d['roofline_lift']['frac'], d['roofline_lift']['ms'])
print('cpu', d.get('cpu_baseline')); print('cuda_ref', d.get('cuda_reference')); print('tm', json.dumps(d.get('throughput_mode'))[:300]); print('clocks', d['clocks'])
r=json.load(open('gpurun_out/r02e_bench_ref.json')); print('ref', r['value'], r['ms_per_step'])
PY
