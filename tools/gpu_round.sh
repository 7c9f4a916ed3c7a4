#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu full captures of the step's kernels, secondary benches.
# usage (on the box, via gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1; cat gpurun_out/bench_ref_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu_$tag.log 2>&1
for spec in "a1 trace wf_forward" "a2 trace wf_forward" "a1 score wf_forward" "a1 trace traceback"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o gpurun_out/prof_${tag}_$1_$2_$3 \
      python tools/profile_forward.py 100000 $1 $2 3 > gpurun_out/ncu_${tag}_$1_$2_$3.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:umi_neighbors -s 2 -c 1 -f -o gpurun_out/prof_${tag}_umi \
    python tools/bench_umi.py 100000 100000 > gpurun_out/ncu_${tag}_umi.log 2>&1
python tools/profile_forward.py 200000 a1 trace 3 | tail -1; python tools/profile_forward.py 200000 a2 trace 3 | tail -1
python tools/profile_forward.py 200000 a1 score 3 | tail -1; python tools/profile_forward.py 200000 a2 score 3 | tail -1
python tools/bench_extra.py > gpurun_out/bench_extra_$tag.log 2>&1; cat gpurun_out/bench_extra_$tag.log
SARLACC_DEBUG_TIMING=1 python tools/bench_umi.py 400000 2000 > gpurun_out/bench_umi_$tag.log 2>&1; SARLACC_DEBUG_TIMING=1 python tools/bench_umi.py 100000 100000 >> gpurun_out/bench_umi_$tag.log 2>&1; cat gpurun_out/bench_umi_$tag.log
