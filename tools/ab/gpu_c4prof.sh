#!/bin/bash
mkdir -p /tmp/prof
for solo in 0 1; do
  SARLACC_SOLO_SHORT=$solo timeout 600 ncu --set full --clock-control none --import-source on -k regex:wf_forward -s 3 -c 1 -f -o /tmp/prof/c4_$solo \
      python tools/bench_c4.py 400000 > gpurun_out/ncu_c4_$solo.log 2>&1
  python tools/ncu_summary.py /tmp/prof/c4_$solo.ncu-rep > gpurun_out/ncu_summary_c4_solo$solo.txt 2>&1
  ncu -i /tmp/prof/c4_$solo.ncu-rep --page raw --csv > gpurun_out/ncu_raw_c4_solo$solo.csv 2>/dev/null
  SARLACC_SOLO_SHORT=$solo timeout 300 python tools/bench_c4.py 1000000 | tail -1
done
cat gpurun_out/ncu_summary_c4_solo0.txt gpurun_out/ncu_summary_c4_solo1.txt
