exec > gpurun_out/run.log 2>&1
python -m pytest tests/test_gpu_umi.py -x -q 2>&1 | tail -15
SARLACC_DEBUG_TIMING=1 python tools/bench_umi.py 400000 2000 2>&1 | tail -12
SARLACC_DEBUG_TIMING=1 python tools/bench_umi.py 100000 100000 2>&1 | grep -v reference | tail -6
