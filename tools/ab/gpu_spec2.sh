#!/bin/bash
SARLACC_DEBUG_SPEC=1 timeout 600 python bench.py --no-cpu --no-extra --no-e2e --steps 2 --warmup 1 2>&1 | grep speculation | head -3
for sp in 1 0; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/spec_launches_$sp.csv env SARLACC_SPECULATE=$sp python bench.py --no-cpu --no-extra --no-e2e --steps 1 --warmup 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/spec_launches_$sp.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
L=[(r[ki][:70], float(r[vi].replace(',',''))/(1e3 if r[ui]=='ns' else 1), ) for r in rows[1:]]
print("SPEC=$sp n launches", len(L))
for k,v in L[-45:]: print("%9.1f us  %s"%(v,k))
PY
done
