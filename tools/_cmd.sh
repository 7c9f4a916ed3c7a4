exec > gpurun_out/run10.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "--- pinned"; PROBE_PINNED=1 SARLACC_DEBUG_TIMING=1 python tools/e2e_probe.py 1000000 3 2>&1 | tail -4
echo "--- pageable, 16 threads"; python tools/e2e_probe.py 1000000 3 2>&1 | tail -1
