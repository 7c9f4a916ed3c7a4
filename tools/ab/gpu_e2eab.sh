#!/bin/bash
for cfg in "X=1" "SARLACC_NO_UPLOAD_CHAIN=1" "SARLACC_UNGATE_CHUNKS=1" "SARLACC_SLOTS=4 SARLACC_UNGATE_CHUNKS=1"; do
  echo "--- $cfg"
  env $cfg timeout 600 python bench.py --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pageable_inputs_reads_per_s'], d['e2e']['phases_ms_per_rank'])"
done
SARLACC_DEBUG_TIMING=1 timeout 600 python bench.py --no-cpu --no-extra --steps 2 --warmup 1 2>&1 >/dev/null | grep sarlacc | sed -n 12,22p
