#!/bin/bash
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -x -m gpu 2>&1 | tail -2
for v in "X=1" "SARLACC_NO_SERPENTINE=1" "SARLACC_GATE_CHUNKS=1" "SARLACC_SOLO_SHORT=0" "SARLACC_CHUNK=284160" "SARLACC_CHUNK=568320"; do
  echo "--- $v"; env $v timeout 300 python tools/bench_c4.py 1000000 | tail -1
done
