#!/bin/bash
# round 2, call I (2 GPUs): whole suite, N=2 configs[4] share with the gather, FASTQ ingest incl. gzip, bench at N=1
tag=${1:-r2i}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$tag.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/run_c5.py --total 50000000 --check-stride 3000017 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3
timeout 600 python tools/bench_fastq.py 100000 5000 2>&1 | tail -8
timeout 1200 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2i.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gcups','step_roofline_frac','gpu_launches')})
print({k:d['e2e'][k] for k in ('value','ms_per_step','unfused_reads_per_s','pageable_inputs_reads_per_s')})
print(d['roofline']['achieved'], d['roofline']['frac'], d['cpu_baseline']['value'])
for k in ('c3','c4','c5'):
    print(k, {x:d[k][x] for x in d[k] if x not in ('workload','cpu_baseline','kernels','phases_ms_rank0')})
PY
