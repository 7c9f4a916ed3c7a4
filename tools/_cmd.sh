exec > gpurun_out/run11.log 2>&1
python -m pytest tests/test_gpu_api.py -x -q -k "pinned or errors" 2>&1 | tail -8
