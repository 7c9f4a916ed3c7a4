#!/bin/bash
# round 2, call C: chunk engine tests + whole suite + kernel rates + configs[4] share
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chunk.py -x -q > gpurun_out/pytest_chunk_$tag.log 2>&1; echo "chunk pytest rc=$?"; tail -15 gpurun_out/pytest_chunk_$tag.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_chunk.py > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
for a in a1 a2; do for m in trace score; do timeout 300 python tools/profile_forward.py 200000 $a $m 3 | tail -1; done; done
timeout 600 python tools/run_c5.py --total 5000000 --check-stride 50021 2>&1 | tail -3
timeout 900 python bench.py --c5-reads 10000000 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
