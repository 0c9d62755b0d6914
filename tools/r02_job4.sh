#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_ops.py tests/test_gpu_sfa.py -q -m gpu ) > gpurun_out/r02_t_kernels.txt 2>&1; echo "kernels rc=$?"; tail -8 gpurun_out/r02_t_kernels.txt
( timeout 900 python -m pytest tests/test_gpu_unet3d.py tests/test_gpu_net2d.py tests/test_gpu_golden.py tests/test_gpu_slab.py tests/test_gpu_dropin.py tests/test_gpu_zz_widening.py -q -m gpu ) > gpurun_out/r02_t_models.txt 2>&1; echo "models rc=$?"; tail -8 gpurun_out/r02_t_models.txt
( timeout 900 python bench.py --steps 10 --warmup 3 --dump-profile --no-cpu ) > gpurun_out/r02_bench_tf32.json 2> gpurun_out/r02_bench_tf32.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_tf32.json; tail -5 gpurun_out/r02_bench_tf32.err
( OCCDEPTH_PRECISION=tf32 timeout 300 python tools/conv_bench.py ) > gpurun_out/r02_convbench_tf32.txt 2>&1
( OCCDEPTH_PRECISION=bf16 timeout 300 python tools/conv_bench.py ) > gpurun_out/r02_convbench_bf16.txt 2>&1
for s in b1_expand bneck5_16_64; do ( OCCDEPTH_PRECISION=tf32 timeout 120 python tools/conv_trace.py $s ) > gpurun_out/r02_trace_tf32_$s.txt 2>&1; done
cat gpurun_out/r02_trace_tf32_b1_expand.txt
