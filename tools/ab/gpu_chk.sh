#!/bin/bash
timeout 1800 python -m pytest tests/test_gpu_chunk.py tests/test_golden_c1.py tests/test_gpu_parity.py -q -x -m gpu 2>&1 | tail -2
timeout 900 python bench.py --no-cpu > gpurun_out/bench_chk.json 2>gpurun_out/bench_chk.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_chk.json') if l.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], d['step_roofline_frac'], 'e2e', d['e2e']['value'], 'c3', d['c3']['ms_core'], d['c3']['roofline_frac_core'], 'c4', d['c4']['seconds_e2e'], d['c4']['roofline_frac_e2e'], 'c5', d['c5']['seconds'], d['c5']['roofline_frac_per_gpu'], d['c5']['parity'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:classify -c 12 --csv --log-file gpurun_out/classify.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra > /dev/null 2>&1
grep classify gpurun_out/classify.csv | tail -4 | cut -d, -f5,12-
