#!/bin/bash
mkdir -p gpurun_out
for pr in tf32 bf16; do for nv in 4 2; do for mb in 1 4; do echo -n "C=64 $pr NV=$nv MINB=$mb: "; ( OCCD_LIFT_NV=$nv OCCD_LIFT_MINB=$mb OCCDEPTH_PRECISION=$pr timeout 100 python tools/lift_bench.py ) 2>&1 | tail -1; done; done; done
( OCCDEPTH_PRECISION=tf32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sfa_lift -s 4 -c 1 -o gpurun_out/r02b_prof_lift -f python tools/lift_bench.py ) > gpurun_out/r02b_ncu_lift.log 2>&1; echo "ncu rc=$?"
