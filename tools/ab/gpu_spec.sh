#!/bin/bash
timeout 1800 python -m pytest tests/test_gpu_chunk.py tests/test_golden_c1.py tests/test_gpu_api.py -q -x -m gpu 2>&1 | tail -3
for cfg in "SARLACC_OVERLAP=1" "SARLACC_GATE_CHUNKS=1" "SARLACC_OVERLAP=0"; do
  echo "--- $cfg"
  env $cfg timeout 600 python bench.py --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('value', d['value'], d['ms_per_step'], d['step_roofline_frac'], 'kernel', d['roofline']['achieved'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"
  env $cfg timeout 600 python tools/run_c5.py --total 10000000 --check-stride 500009 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('c5', d['seconds'], d['gcups_per_gpu'], d['phases_ms_rank0'], d['parity'])"
done
