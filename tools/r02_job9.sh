#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for pr in tf32 bf16; do
  ( timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_ncu_$pr.csv python tools/ncu_forward.py $pr ) > gpurun_out/r02_ncu_$pr.log 2>&1; echo "ncu $pr rc=$?"; tail -2 gpurun_out/r02_ncu_$pr.log
done
# one --set full capture each of the three dominant kernels (tf32): x-packed halo head conv, per-tap decoder conv, lift
( OCCDEPTH_PRECISION=tf32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halox -s 4 -c 1 -o gpurun_out/r02_prof_halox_head -f python tools/conv_bench.py head_c32_d1 ) > gpurun_out/r02_ncu_full1.log 2>&1; echo "full1 rc=$?"
( OCCDEPTH_PRECISION=tf32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 4 -c 1 -o gpurun_out/r02_prof_tc_up2 -f python tools/conv_bench.py up2_conv2 ) > gpurun_out/r02_ncu_full2.log 2>&1; echo "full2 rc=$?"
( OCCDEPTH_PRECISION=tf32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sfa_lift -s 4 -c 1 -o gpurun_out/r02_prof_lift -f python tools/lift_bench.py ) > gpurun_out/r02_ncu_full3.log 2>&1; echo "full3 rc=$?"
( OCCDEPTH_PRECISION=tf32 timeout 100 python tools/lift_bench.py ) > gpurun_out/r02_liftbench_tf32.txt 2>&1; cat gpurun_out/r02_liftbench_tf32.txt
( timeout 900 python -m pytest tests/test_gpu_net2d.py tests/test_gpu_config2.py -q -m gpu -k "infer_mode or flosp_depth_vs" -s ) > gpurun_out/r02_t_new.txt 2>&1; echo "new tests rc=$?"; tail -12 gpurun_out/r02_t_new.txt
ls -la gpurun_out/*.ncu-rep
