#!/bin/bash
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -x -m gpu 2>&1 | tail -2
for cfg in "X=1" "SARLACC_NO_UPLOAD_CHAIN=1"; do
echo "--- $cfg"
env $cfg timeout 900 python bench.py --no-cpu --no-extra 2>/dev/null | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'pageable', d['e2e']['pageable_inputs_reads_per_s'], 'unfused', d['e2e']['unfused_reads_per_s'])"
env $cfg timeout 300 python tools/bench_c4.py 1000000 | tail -1
done
