#!/bin/bash
for i in 1 2 3; do
timeout 600 python tools/run_c5.py --total 10000000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print({k: d[k] for k in ('seconds','seconds_alignment_passes','seconds_gather_rank0','seconds_thresholds_rank0','gcups_per_gpu','phases_ms_rank0')})"
done
