#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02k_bench_n2.json 2> gpurun_out/r02k_bench_n2.err; echo "rc=$? stdout_lines=$(wc -l < gpurun_out/r02k_bench_n2.json)"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench_n2.json').read())
print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['pipelined_error'])
print(json.dumps(d.get('slab'))[:500])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02k_bench_ref_n2.json 2> gpurun_out/r02k_bench_ref_n2.err; echo "ref rc=$? stdout_lines=$(wc -l < gpurun_out/r02k_bench_ref_n2.json)"; head -c 200 gpurun_out/r02k_bench_ref_n2.json
