exec > gpurun_out/run13.log 2>&1
python -m pytest tests -m gpu -x -q -k "full_size or pipeline or resident or tune or thresholds" 2>&1 | tail -3
python tools/run_c5.py --share 1250000 --batch 625000
