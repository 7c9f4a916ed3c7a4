#!/bin/bash
timeout 1800 python -m pytest tests/test_gpu_api.py -q -x -m gpu 2>&1 | tail -2
for cfg in "X=1" "SARLACC_PACK_SEQ=1" "SARLACC_PACK_SEQ=0"; do
echo "--- $cfg"
env $cfg timeout 900 python bench.py --no-cpu --no-extra 2>/dev/null | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']
print('value', d['value'], 'e2e', e['value'], e['ms_per_step'], 'pageable', e['pageable_inputs_reads_per_s'], 'h2d', e['h2d_bytes_per_step'], e['h2d_bytes_per_step_plain'], e['upload_gbs_per_rank'], e['phases_ms_per_rank'])"
done
