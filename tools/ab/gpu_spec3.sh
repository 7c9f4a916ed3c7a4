#!/bin/bash
timeout 1800 python -m pytest tests/test_gpu_chunk.py tests/test_golden_c1.py tests/test_gpu_api.py -q -x -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_c4.py 1000000
SARLACC_DEBUG_TIMING=1 timeout 300 python tools/bench_c4.py 1000000 2>&1 | tail -12
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c4_launches.csv python tools/bench_c4.py 1000000 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/c4_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
L=[(r[ki][:90], float(r[vi].replace(',',''))/(1e3 if r[ui]=='ns' else 1), ) for r in rows[1:]]
print("n launches", len(L))
for k,v in L[-40:]: print("%9.1f us  %s"%(v,k))
PY
