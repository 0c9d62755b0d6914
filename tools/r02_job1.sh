#!/bin/bash
# round-2 GPU job 1: peaks probe, experimental variant parity + A/B, role traces, per-shape conv bench per impl
mkdir -p gpurun_out
( timeout 300 python tools/peaks_probe.py ) > gpurun_out/r02_peaks_probe.log 2>&1; echo "peaks rc=$?"
AB_PROFILE=1 bash tools/validate_variants.sh > gpurun_out/r02_variants.log 2>&1; echo "variants rc=$?"
for s in b1_expand bneck5_16_64 expand_188 l1_k113_16; do ( timeout 120 python tools/conv_trace.py $s ) > gpurun_out/r02_trace_$s.txt 2>&1; done
( timeout 300 python tools/conv_bench.py ) > gpurun_out/r02_convbench_default.txt 2>&1
( BENCH_IMPL=halox timeout 200 python tools/conv_bench.py head_c32_d1 head_c32_d2 head_c32_d3 head_c32_n2 x_c32_n32 x_c64_n32 x_c32_n64 x_c64_n64 x_c16_n16 l1_k113_16 ) > gpurun_out/r02_convbench_halox.txt 2>&1
( BENCH_IMPL=tcx timeout 200 python tools/conv_bench.py up1_conv2 head_c32_d1 x_c64_n64 ) > gpurun_out/r02_convbench_tcx.txt 2>&1
( BENCH_IMPL=tcm2 timeout 200 python tools/conv_bench.py up2_conv2 up4_conv2 up16_conv1 x_c64_n256 ) > gpurun_out/r02_convbench_tcm2.txt 2>&1
tail -5 gpurun_out/r02_peaks_probe.log; cat gpurun_out/variants_ab.txt; cat gpurun_out/r02_convbench_halox.txt
