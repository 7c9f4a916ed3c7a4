#!/bin/bash
# bench.py under torchrun at N GPUs (default 8), as the driver launches it
N=${1:-8}; tag=${2:-n8}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo_$tag.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$tag.json') if x.startswith('{')][-1]
d=json.loads(l)
print({k:d[k] for k in ('value','ms_per_step','gcups','step_roofline_frac','n_gpus')})
e=d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['upload_gbs_per_rank']); print(e['phases_ms_per_rank'])
for k in ('c3','c4','c5'): print(k,{x:d[k][x] for x in d[k] if x not in ('workload','cpu_baseline','kernels')})
PY
cat gpurun_out/topo_$tag.txt | cut -c1-160
