#!/bin/bash
for v in main ss main ss; do
  echo "--- $v"
  if [ $v = main ]; then L=""; else L="variants/lib_$v.so"; fi
  SARLACC_LIB=$L timeout 300 python tools/profile_forward.py 400000 a1 score 4 | tail -2
done
SARLACC_LIB=variants/lib_ss.so timeout 600 python bench.py --no-cpu --no-e2e --c5-reads 5000000 --c4-sequences 200000 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('ss value', d['value'], d['ms_per_step'], d['step_roofline_frac'], 'c3', d['c3'].get('ms_core'), d['c3'].get('roofline_frac_core'), 'c5', d['c5'].get('seconds'), d['c5'].get('gcups_per_gpu'))"
timeout 600 python bench.py --no-cpu --no-e2e --c5-reads 5000000 --c4-sequences 200000 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('main value', d['value'], d['ms_per_step'], d['step_roofline_frac'], 'c3', d['c3'].get('ms_core'), d['c3'].get('roofline_frac_core'), 'c5', d['c5'].get('seconds'), d['c5'].get('gcups_per_gpu'))"
