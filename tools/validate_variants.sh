#!/bin/bash
# One gpurun call that (1) runs the opt-in kernel variants' parity cases, (2) A/Bs them on the config-2 forward.
#   gpurun --timeout 1500 -- 'bash tools/validate_variants.sh'
# Every step runs under its own timeout; results land in gpurun_out/variants_*.txt and ab_experiments.json.
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; ( "$@" ) > gpurun_out/variants_$name.txt 2>&1; echo "rc=$? ($name)"; tail -3 gpurun_out/variants_$name.txt; }
run tests_experimental env OCCD_EXPERIMENTAL=1 timeout 420 python -m pytest tests/test_gpu_conv.py tests/test_gpu_ops.py -q -m gpu
run tests_epiwide env OCCD_EPI_WIDE=1 timeout 240 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet3d.py -q -m gpu
run tests_pdl env OCCD_PDL=1 timeout 240 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet3d.py tests/test_gpu_net2d.py -q -m gpu
run tests_stages env OCCD_TC_STAGES_MIN=16 timeout 240 python -m pytest tests/test_gpu_conv.py -q -m gpu
run ab timeout 900 python tools/ab_experiments.py "$@"
cat gpurun_out/variants_ab.txt
