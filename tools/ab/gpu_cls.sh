#!/bin/bash
mkdir -p /tmp/prof
timeout 600 ncu --set full --clock-control none -k regex:"classify_strands" -s 2 -c 2 -f -o /tmp/prof/cls python tools/run_c5.py --total 454656 --chunk 454656 > gpurun_out/ncu_cls.log 2>&1
python tools/ncu_summary.py /tmp/prof/cls.ncu-rep > gpurun_out/ncu_summary_classify.txt 2>&1
cat gpurun_out/ncu_summary_classify.txt | head -40
