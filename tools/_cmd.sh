exec > gpurun_out/run2.log 2>&1
python -m pytest tests/test_gpu_api.py -x -q -k "pipeline or sharding or fused" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1
nproc
python tools/bench_extra.py 2>&1 | head -2
