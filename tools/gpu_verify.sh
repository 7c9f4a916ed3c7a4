#!/bin/bash
# Short GPU pass: the whole -m gpu suite, smoke, both bench arms (no profiles).  usage: bash tools/gpu_verify.sh <tag>
tag=${1:-v}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_$tag.json') if l.startswith('{')][-1])
e=d['e2e']
print('value', d['value'], d['ms_per_step'], d['step_roofline_frac'], 'kernel', d['roofline']['achieved'], d['roofline']['frac'])
print('e2e', e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['pageable_inputs_reads_per_s'], e['unfused_reads_per_s'])
print('c3', d['c3']['ms_core'], d['c3']['roofline_frac_core'], 'c4', d['c4']['seconds_e2e'], d['c4']['roofline_frac_e2e'], 'c5', d['c5']['seconds'], d['c5']['roofline_frac_per_gpu'], d['c5']['parity'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
