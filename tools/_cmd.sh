python -m pytest tests -m gpu -x -q 2>&1 | tail -2
SARLACC_DEBUG_TIMING=1 python tools/e2e_probe.py 1000000 3 2>&1 | tail -3
python bench.py --no-cpu
