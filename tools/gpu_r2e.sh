#!/bin/bash
# round 2, call E: generator parity after the rewrite, kernel A/B (bank-spread tables + variants), configs[4] share
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chunk.py -q -x > gpurun_out/pytest_chunk_$tag.log 2>&1; echo "chunk pytest rc=$?"; tail -4 gpurun_out/pytest_chunk_$tag.log
echo "--- main build"
for a in a1 a2; do for m in trace score; do timeout 300 python tools/profile_forward.py 200000 $a $m 3 | tail -1; done; done
for v in short solo2 xl2; do
  echo "--- variant $v"
  for a in a1 a2; do for m in trace score; do
    if [ $v = xl2 ] && [ $a = a1 ]; then continue; fi
    SARLACC_LIB=variants/lib_$v.so timeout 300 python tools/profile_forward.py 200000 $a $m 3 | tail -1; done; done
done
timeout 600 python tools/run_c5.py --total 10000000 --check-stride 200003 2>&1 | tail -2
