exec > gpurun_out/run8b.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "--- pinned, 16 threads"; PROBE_PINNED=1 SARLACC_DEBUG_TIMING=1 python tools/e2e_probe.py 1000000 3 2>&1 | tail -12
echo "--- pinned, 2 threads"; PROBE_PINNED=1 SARLACC_HOST_THREADS=2 python tools/e2e_probe.py 1000000 3 2>&1 | tail -1
echo "--- pageable, 16 threads"; SARLACC_DEBUG_TIMING=1 python tools/e2e_probe.py 1000000 3 2>&1 | tail -2
echo "--- pageable, 4 threads"; SARLACC_HOST_THREADS=4 python tools/e2e_probe.py 1000000 3 2>&1 | tail -1
echo "--- pageable, 4 threads, host pack"; SARLACC_HOST_PACK=1 SARLACC_HOST_THREADS=4 python tools/e2e_probe.py 1000000 3 2>&1 | tail -1
