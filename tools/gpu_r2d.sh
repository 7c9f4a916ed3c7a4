#!/bin/bash
# round 2, call D: chunk tests, threshold timing, ncu captures of the forward kernels / traceback / generator (CSV exports
# are made on the box: the .ncu-rep files of a --set full capture with source are ~17 MB each)
tag=${1:-r2d}
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests/test_gpu_chunk.py -q > gpurun_out/pytest_chunk_$tag.log 2>&1; echo "chunk pytest rc=$?"; tail -4 gpurun_out/pytest_chunk_$tag.log
SARLACC_DEBUG_TIMING=1 timeout 600 python tools/run_c5.py --total 10000000 2>&1 | grep -v "chunk at\|pair job" | tail -8
for spec in "a1 trace" "a1 score" "a2 trace" "a2 score"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:wf_forward -s 2 -c 1 -f -o /tmp/prof/${1}_$2 \
      python tools/profile_forward.py 100000 $1 $2 3 > gpurun_out/ncu_${tag}_$1_$2.log 2>&1
  ncu -i /tmp/prof/${1}_$2.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_$1_$2_raw.csv 2>/dev/null
  ncu -i /tmp/prof/${1}_$2.ncu-rep --page source --csv > gpurun_out/ncu_${tag}_$1_$2_source.csv 2>/dev/null
done
cp /tmp/prof/a1_trace.ncu-rep gpurun_out/prof_${tag}_a1_trace.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"traceback|mock_windows|mock_widths|scramble_rows|resolve_strand" -c 8 -f -o /tmp/prof/aux \
    python tools/run_c5.py --total 200000 --chunk 200000 > gpurun_out/ncu_${tag}_aux.log 2>&1
ncu -i /tmp/prof/aux.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_aux_raw.csv 2>/dev/null
du -sh gpurun_out
