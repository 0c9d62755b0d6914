#!/bin/bash
# round-2 GPU job 2: first run of the TF32 mode -- kernel tests, model tests, full-size parity, bench + profile
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_ops.py tests/test_gpu_sfa.py -q -m gpu -x ) > gpurun_out/r02_t_kernels.txt 2>&1; echo "kernels rc=$?"; tail -5 gpurun_out/r02_t_kernels.txt
( timeout 900 python -m pytest tests/test_gpu_unet3d.py tests/test_gpu_net2d.py tests/test_gpu_golden.py tests/test_gpu_slab.py tests/test_gpu_zz_widening.py -q -m gpu ) > gpurun_out/r02_t_models.txt 2>&1; echo "models rc=$?"; tail -15 gpurun_out/r02_t_models.txt
( timeout 900 python -m pytest tests/test_gpu_config2.py tests/test_gpu_config4.py -q -m gpu -s ) > gpurun_out/r02_t_configs.txt 2>&1; echo "configs rc=$?"; tail -12 gpurun_out/r02_t_configs.txt
( timeout 900 python bench.py --steps 10 --warmup 3 --dump-profile ) > gpurun_out/r02_bench_tf32.json 2> gpurun_out/r02_bench_tf32.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r02_bench_tf32.json; tail -5 gpurun_out/r02_bench_tf32.err
( OCCDEPTH_PRECISION=tf32 timeout 300 python tools/conv_bench.py ) > gpurun_out/r02_convbench_tf32.txt 2>&1
( timeout 120 python __graft_entry__.py smoke ) > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r02_smoke.txt
cat gpurun_out/parity_measured.jsonl
