#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "rc=$?"
python - <<'PY'
import json
t=open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()
print('stdout lines', len(t))
d=json.loads(t[-1])
print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], d['step_ms'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['pipelined_error'])
print(json.dumps(d.get('slab'))[:700])
PY
tail -3 gpurun_out/r02_bench_n8.err
