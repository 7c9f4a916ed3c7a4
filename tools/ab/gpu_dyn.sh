#!/bin/bash
for d in 1 0; do
  echo "--- SARLACC_DYNAMIC=$d"
  for a in a1 a2; do for m in trace score; do SARLACC_DYNAMIC=$d timeout 300 python tools/profile_forward.py 200000 $a $m 3 | tail -1; done; done
  SARLACC_DYNAMIC=$d timeout 600 python bench.py --no-cpu --no-e2e --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['step_roofline_frac'])"
done
SARLACC_DYNAMIC=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chunk.py -q -x 2>&1 | tail -2
