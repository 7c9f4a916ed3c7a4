exec > gpurun_out/run5.log 2>&1
for v in J3 J4; do echo $v; for m in trace score; do SARLACC_LIB=scratch/lib_$v.so python tools/profile_forward.py 200000 a1 $m 3 | tail -1; done; done
