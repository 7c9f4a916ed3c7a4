#!/bin/bash
for cfg in "X=1" "SARLACC_CHUNK=170496" "SARLACC_CHUNK=227328 SARLACC_SCRATCH_MB=8192" "SARLACC_CHUNK=56832"; do
echo "--- $cfg"
env $cfg timeout 600 python bench.py --no-cpu --no-extra --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], d['step_roofline_frac'], d['gpu_launches'])"
done
