#!/bin/bash
# round 2, call G: whole suite with both oracles + new tests; configs[3] A/B
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$tag.log
timeout 300 python tools/bench_c4.py
SARLACC_NO_SOLO=1 timeout 300 python tools/bench_c4.py
SARLACC_NO_LENGTH_ORDER=1 timeout 300 python tools/bench_c4.py
timeout 600 ncu --set full --clock-control none -k regex:wf_forward -s 1 -c 1 -f -o /tmp/c4 python tools/bench_c4.py 200000 > gpurun_out/ncu_${tag}_c4.log 2>&1
ncu -i /tmp/c4.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_c4_raw.csv 2>/dev/null
