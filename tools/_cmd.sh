exec > gpurun_out/run14.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/run_c5.py --share 1250000 --batch 625000
python bench.py --no-cpu | cut -c1-1500
