#!/bin/bash
mkdir -p gpurun_out
for cfg in "20 5" "10 3" "20 5" "50 10"; do set -- $cfg; ( timeout 600 python bench.py --steps $1 --warmup $2 --no-cpu --no-modes ) 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', d['ms_per_step'], d['step_ms'], d['e2e']['ms_per_step'])"; done
