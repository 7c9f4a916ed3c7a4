exec > gpurun_out/run.log 2>&1
echo default; python tools/bench_extra.py 2>&1 | head -2
echo again; python tools/bench_extra.py 2>&1 | head -1
echo CHUNK131072; SARLACC_CHUNK=131072 python tools/bench_extra.py 2>&1 | head -1
echo NOAVX; SARLACC_NO_AVX2=1 python tools/bench_extra.py 2>&1 | head -1
echo PAIR0; SARLACC_PAIR=0 python tools/bench_extra.py 2>&1 | head -1
echo THREADS8; SARLACC_HOST_THREADS=8 python tools/bench_extra.py 2>&1 | head -1
