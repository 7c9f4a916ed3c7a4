python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SARLACC_DEBUG_TIMING=1 python tools/e2e_probe.py 1000000 3 2>&1 | tail -8
SARLACC_DEBUG_TIMING=1 SARLACC_HOST_THREADS=8 python tools/e2e_probe.py 1000000 2 2>&1 | tail -5
SARLACC_DEBUG_TIMING=1 SARLACC_CHUNK=65536 python tools/e2e_probe.py 1000000 2 2>&1 | tail -5
python bench.py --no-cpu --no-e2e
SARLACC_NO_ENDROW=1 python bench.py --no-cpu --no-e2e
nproc; lscpu | grep -i "model name\|socket\|numa" 
