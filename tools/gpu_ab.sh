#!/bin/bash
# A/B of variant libraries on the forward kernels: bash tools/gpu_ab.sh "v1 v2 ..." "a1 trace|a2 trace|..."
for v in main $1; do
  echo "--- $v"
  for spec in "a1 trace" "a2 trace"; do
    set -- $spec
    if [ $v = main ]; then timeout 300 python tools/profile_forward.py 200000 $1 $2 3 | tail -1
    else SARLACC_LIB=variants/lib_$v.so timeout 300 python tools/profile_forward.py 200000 $1 $2 3 | tail -1; fi
  done
done
