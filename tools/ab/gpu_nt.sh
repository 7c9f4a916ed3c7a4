#!/bin/bash
for cfg in "X=1" "SARLACC_NO_NT_COPY=1" "X=2"; do
echo "--- $cfg"
env $cfg timeout 900 python bench.py --no-cpu --no-extra 2>/dev/null | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']
print('e2e', e['value'], e['ms_per_step'], 'pageable', e['pageable_inputs_reads_per_s'], 'unfused', e['unfused_reads_per_s'])"
done
timeout 300 python -m pytest tests/test_gpu_api.py -q -x -m gpu 2>&1 | tail -1
