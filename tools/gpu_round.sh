#!/bin/bash
# One GPU-box pass (1 GPU): parity tests, smoke, both bench arms, ncu launch list, ncu full captures of the step's kernels
# (summaries made on the box; one .ncu-rep brought back), secondary benches, sanitizers.
# usage (on the box, via gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out /tmp/prof
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$tag.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1; cut -c1-300 gpurun_out/bench_ref_$tag.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/bench_under_ncu_$tag.log 2>&1
for spec in "a1 trace" "a1 score" "a2 trace" "a2 score"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:wf_forward -s 2 -c 1 -f -o /tmp/prof/${1}_$2 \
      python tools/profile_forward.py 100000 $1 $2 3 > gpurun_out/ncu_${tag}_$1_$2.log 2>&1
  python tools/ncu_summary.py /tmp/prof/${1}_$2.ncu-rep $(( 100000 * 250 * $( [ $1 = a1 ] && echo 70 || echo 22 ) )) > gpurun_out/ncu_summary_${tag}_$1_$2.txt 2>&1
done
cp /tmp/prof/a1_trace.ncu-rep gpurun_out/prof_${tag}_a1_trace.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"traceback|mock_windows|scramble_rows|resolve_strand|pack_rows|classify_strands" -c 12 -f -o /tmp/prof/aux \
    python tools/run_c5.py --total 200000 --chunk 200000 > gpurun_out/ncu_${tag}_aux.log 2>&1
python tools/ncu_summary.py /tmp/prof/aux.ncu-rep > gpurun_out/ncu_summary_${tag}_aux.txt 2>&1
for a in a1 a2; do for m in trace score; do python tools/profile_forward.py 200000 $a $m 3 | tail -1; done; done
timeout 300 python tools/bench_c4.py 2>&1 | tail -1
timeout 600 python tools/bench_fastq.py 200000 5000 > gpurun_out/bench_fastq_$tag.log 2>&1; cat gpurun_out/bench_fastq_$tag.log
SARLACC_DEBUG_TIMING=1 timeout 600 python tools/bench_umi.py 400000 2000 > gpurun_out/bench_umi_$tag.log 2>&1; tail -4 gpurun_out/bench_umi_$tag.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py 200 > gpurun_out/sanitizer_memcheck_$tag.log 2>&1; echo "memcheck rc=$?"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py 120 > gpurun_out/sanitizer_racecheck_$tag.log 2>&1; echo "racecheck rc=$?"
du -sh gpurun_out
