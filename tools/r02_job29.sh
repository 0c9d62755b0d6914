#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
( timeout 2400 python -m pytest tests -q -m gpu -x ) > gpurun_out/r02_final_gpu_tests.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_final_gpu_tests.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -3
