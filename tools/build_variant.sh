#!/bin/bash
# Builds scratch/lib_<tag>.so = the library with extra -D flags on kernels.cu (A/B runs: SARLACC_LIB=scratch/lib_<tag>.so).
# usage: tools/build_variant.sh tag "-DSARLACC_WF_BLOCKS_LARGE=3 ..."
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p scratch
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-pthread -Iinclude -Isarlacc_b200/csrc"
nvcc $F -Xptxas -v $@ -c -o scratch/kernels_$tag.o sarlacc_b200/csrc/kernels.cu 2> scratch/$tag.log
nvcc -shared -o scratch/lib_$tag.so scratch/kernels_$tag.o sarlacc_b200/csrc/api.o sarlacc_b200/csrc/umi.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -pthread
grep -A3 "wf_forward2ILi1[18]ELb1" scratch/$tag.log | grep -E "spill|Used"
