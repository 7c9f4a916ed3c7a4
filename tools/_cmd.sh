SARLACC_DEBUG_TIMING=1 python tools/e2e_probe.py 1000000 2 2>&1 | tail -14
