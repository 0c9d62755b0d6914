#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-modes ) > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; tail -c 900 gpurun_out/r02_bench_n2.json; tail -6 gpurun_out/r02_bench_n2.err
for g in 0 4 16; do ( OCCD_LIFT_SHAPE=$g OCCDEPTH_PRECISION=tf32 timeout 100 python tools/lift_bench.py ) 2>&1 | tail -1; done
( timeout 600 python -m pytest tests/test_gpu_net2d.py -q -m gpu -k "infer_mode" ) 2>&1 | tail -5
