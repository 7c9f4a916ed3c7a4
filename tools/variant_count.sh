#!/bin/bash
# static per-cell instruction count of one geometry built with extra -D flags (no GPU): tools/variant_count.sh C solo "flags"
C=$1; solo=$2; shift 2
out=scratch/var_$$.cubin
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -std=c++17 -Iinclude -Isarlacc_b200/csrc -DSARLACC_ONLY_C=$C "$@" -Xptxas -v -cubin -o $out sarlacc_b200/csrc/kernels.cu 2> scratch/var_$$.log || { tail -5 scratch/var_$$.log; exit 1; }
for t in 1 0; do python tools/sass_loop.py $C $t wf_forward2 $solo --lib $out | head -1; done
grep -A2 "wf_forward2ILi${C}E" scratch/var_$$.log | grep -E "spill" | tr '\n' ' '; echo
rm -f $out scratch/var_$$.log
