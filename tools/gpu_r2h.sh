#!/bin/bash
# round 2, call H (2 GPUs): bench under torchrun at N=2 (c5 with the gather), then sanitizers on GPU 0
tag=${1:-r2h}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err; echo "bench N=2 rc=$?"; cat gpurun_out/bench_n2_$tag.json; tail -3 gpurun_out/bench_n2_$tag.err
timeout 900 python tools/sanitize_run.py 300 2>&1 | tail -2
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py 200 > gpurun_out/sanitizer_memcheck_$tag.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_$tag.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py 120 > gpurun_out/sanitizer_racecheck_$tag.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_$tag.log
