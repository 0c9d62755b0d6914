#!/bin/bash
mkdir -p gpurun_out
( timeout 60 tools/bin/tma_store_test ) > gpurun_out/r02_tma_store_test.txt 2>&1; echo "tma rc=$?"; cat gpurun_out/r02_tma_store_test.txt | tail -30
( timeout 120 tools/bin/mma_issue_bench ) > gpurun_out/r02_mma_issue_bench.txt 2>&1; echo "mma rc=$?"; cat gpurun_out/r02_mma_issue_bench.txt
( timeout 300 python tools/timing_diag.py tf32 ) > gpurun_out/r02_timing_diag.txt 2>&1; echo "diag rc=$?"; tail -15 gpurun_out/r02_timing_diag.txt
