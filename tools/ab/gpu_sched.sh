#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_api.py -q -x -m gpu 2>&1 | tail -1
for i in 1 2; do
timeout 900 python bench.py --no-cpu --no-extra 2>/dev/null | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']
print('value', d['value'], 'e2e', e['value'], e['ms_per_step'], 'pageable', e['pageable_inputs_reads_per_s'], e['phases_ms_per_rank'])"
done
SARLACC_DEBUG_TIMING=1 timeout 600 python bench.py --no-cpu --no-extra --steps 2 --warmup 1 2>&1 >/dev/null | grep sarlacc | sed -n 14,27p
