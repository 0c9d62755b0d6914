#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "rc=$?"
tail -c 3000 gpurun_out/r02_bench_n4.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'slab', json.dumps(d.get('slab'))[:600])
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n4.json 2> gpurun_out/r02_bench_ref_n4.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r02_bench_ref_n4.json
