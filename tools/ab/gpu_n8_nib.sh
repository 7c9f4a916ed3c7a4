#!/bin/bash
# N = 8: host-buffer path with the upload-bound policy (bases as 4-bit codes) vs plain bytes
N=${1:-8}
for cfg in "X=1" "SARLACC_PACK_SEQ=0" "SARLACC_PACK_SEQ=1"; do
echo "--- $cfg"
env $cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']
print('value', d['value'], 'e2e', e['value'], e['ms_per_step'], 'h2d/rank MB', [round(x/1e6) for x in e['h2d_bytes_per_step_per_rank']], 'GB/s', [round(x,1) for x in e['upload_gbs_per_rank']])
print([ (round(p['host_stage'],1), round(p['host_enqueue'],1), round(p['host_wait_copy_out'],1), round(p['device_upload_sum'],1)) for p in e['phases_ms_per_rank']])"
done
