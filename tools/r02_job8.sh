#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_ops.py tests/test_gpu_sfa.py -q -m gpu -x ) > gpurun_out/r02_t_kernels.txt 2>&1; echo "kernels rc=$?"; tail -4 gpurun_out/r02_t_kernels.txt
( timeout 900 python -m pytest tests/test_gpu_unet3d.py tests/test_gpu_net2d.py tests/test_gpu_golden.py tests/test_gpu_slab.py tests/test_gpu_dropin.py tests/test_gpu_zz_widening.py -q -m gpu ) > gpurun_out/r02_t_models.txt 2>&1; echo "models rc=$?"; tail -4 gpurun_out/r02_t_models.txt
( timeout 900 python bench.py --steps 10 --warmup 3 --dump-profile --no-cpu ) > gpurun_out/r02_bench_tf32.json 2> gpurun_out/r02_bench_tf32.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_tf32.json; tail -5 gpurun_out/r02_bench_tf32.err
( OCCD_PDL=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-modes ) > gpurun_out/r02_bench_tf32_nopdl.json 2>/dev/null; echo "bench nopdl rc=$?"
( OCCDEPTH_PRECISION=tf32 timeout 200 python tools/dw_bench.py ) > gpurun_out/r02_dwbench_tf32.txt 2>&1
( OCCDEPTH_PRECISION=bf16 timeout 200 python tools/dw_bench.py ) > gpurun_out/r02_dwbench_bf16.txt 2>&1
cat gpurun_out/r02_dwbench_tf32.txt gpurun_out/r02_dwbench_bf16.txt
python - <<'PY'
import json
for f in ("r02_bench_tf32.json","r02_bench_tf32_nopdl.json"):
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["profile_ms"])
    except Exception as e: print(f, "ERR", e)
PY
