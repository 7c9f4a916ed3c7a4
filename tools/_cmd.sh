exec > gpurun_out/run3.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_extra.py 2>&1 | head -3
python tools/bench_fastq.py 200000 5000
