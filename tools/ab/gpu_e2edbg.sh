#!/bin/bash
SARLACC_DEBUG_TIMING=1 timeout 600 python bench.py --no-cpu --no-extra --steps 2 --warmup 1 > gpurun_out/e2e_dbg.out 2> gpurun_out/e2e_dbg.err
grep -n "sarlacc" gpurun_out/e2e_dbg.err | tail -60
