#!/bin/bash
# One GPU-box pass: parity tests, smoke, both bench arms, ncu launch list, ncu full captures of the step's kernels,
# secondary benches.  usage (on the box, via gpurun): bash tools/gpu_round.sh <tag> [quick]
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1; cat gpurun_out/bench_ref_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu_$tag.log 2>&1
specs=("a1 trace wf_forward" "a2 trace wf_forward" "a1 trace traceback")
[ "$2" = quick ] || specs+=("a1 score wf_forward")
for spec in "${specs[@]}"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o gpurun_out/prof_${tag}_$1_$2_$3 \
      python tools/profile_forward.py 100000 $1 $2 3 > gpurun_out/ncu_${tag}_$1_$2_$3.log 2>&1
done
for a in a1 a2; do for m in trace score; do python tools/profile_forward.py 200000 $a $m 3 | tail -1; done; done
python tools/bench_extra.py > gpurun_out/bench_extra_$tag.log 2>&1; cat gpurun_out/bench_extra_$tag.log
python tools/run_c5.py --share 1250000 --batch 1250000 > gpurun_out/run_c5_$tag.log 2>&1; cat gpurun_out/run_c5_$tag.log
SARLACC_DEBUG_TIMING=1 python tools/bench_umi.py 400000 2000 > gpurun_out/bench_umi_$tag.log 2>&1; cat gpurun_out/bench_umi_$tag.log
python tools/bench_fastq.py 200000 5000 > gpurun_out/bench_fastq_$tag.log 2>&1; cat gpurun_out/bench_fastq_$tag.log
