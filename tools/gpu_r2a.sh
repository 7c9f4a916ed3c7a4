#!/bin/bash
# round 2, call A: parity suite on the new record layout + solo geometry, A/B of the a2 kernels
tag=${1:-r2b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
for a in a1 a2; do for m in trace score; do timeout 300 python tools/profile_forward.py 200000 $a $m 3 | tail -1; done; done
echo "--- SARLACC_NO_SOLO=1"
for m in trace score; do SARLACC_NO_SOLO=1 timeout 300 python tools/profile_forward.py 200000 a2 $m 3 | tail -1; done
timeout 600 python bench.py --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
