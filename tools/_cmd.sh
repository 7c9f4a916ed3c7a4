exec > gpurun_out/run12.log 2>&1
python tools/run_c5.py --share 2500000 --batch 1250000
