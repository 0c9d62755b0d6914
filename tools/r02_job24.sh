#!/bin/bash
mkdir -p gpurun_out
( OCCDEPTH_PRECISION=tf32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_tiled -s 4 -c 1 -o gpurun_out/r02_prof_dw_b2 -f python tools/dw_bench.py b2_c480_k5 ) > gpurun_out/r02_ncu_dw1.log 2>&1; echo "rc=$?"
( OCCDEPTH_PRECISION=tf32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_tiled -s 4 -c 1 -o gpurun_out/r02_prof_dw_b1 -f python tools/dw_bench.py b1_c288_k3 ) > gpurun_out/r02_ncu_dw2.log 2>&1; echo "rc=$?"
